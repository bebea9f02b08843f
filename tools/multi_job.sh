# usage: bash tools/multi_job.sh N [TAG]  -- on N GPUs of one box: the sharded parity check, the multi-GPU tests, then the driver's
# bench command (default workload + row-sharded sub-record) and the reference arm under torchrun.
N=${1:-2}
TAG=${2:-a}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tools/sharded_check.py > gpurun_out/r02_sharded_check_n${N}.log 2>&1; echo "sharded_check rc=$?"; grep -E "OK|MISMATCH|UNEXPECTED|Error|error" gpurun_out/r02_sharded_check_n${N}.log | tail -40
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg3_n${N}_${TAG}.json 2> gpurun_out/r02_bench_cfg3_n${N}_${TAG}.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_cfg3_n${N}_${TAG}.err
timeout 300 $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n${N}_${TAG}.json 2> gpurun_out/r02_bench_reference_n${N}_${TAG}.err; echo "reference rc=$?"
timeout 300 python -m pytest tests -m gpu -x -q -k "multi" 2>&1 | tail -5

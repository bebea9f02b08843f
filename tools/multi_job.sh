# usage: bash tools/multi_job.sh N   -- cfg4 (bank row-sharded, NCCL all-gather merge) and cfg3 (DP over recordings) at N GPUs
N=${1:-2}
for w in cfg4 cfg3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --no-cpu > gpurun_out/bench_${w}_n${N}.json 2> gpurun_out/bench_${w}_n${N}.err
  tail -c 2500 gpurun_out/bench_${w}_n${N}.json; tail -5 gpurun_out/bench_${w}_n${N}.err
done
timeout 300 python -m pytest tests -m gpu -x -q -k multi 2>&1 | tail -5

# 2 GPUs, the shipped build: sharded parity check (now with a small-query case: bank-stream kernels on every shard) + the multi-GPU CLI tests
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 200 $TR tools/sharded_check.py > gpurun_out/r02_sharded_check_n${N}.log 2>&1; echo "sharded_check rc=$?"; grep -E "OK|MISMATCH|UNEXPECTED|Error|error" gpurun_out/r02_sharded_check_n${N}.log | tail -40
timeout 200 python -m pytest tests -m gpu -x -q --timeout 150 -k "multi" 2>&1 | tail -5 | tee gpurun_out/r02_gputests_multi_n${N}.log

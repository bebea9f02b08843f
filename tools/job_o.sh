timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 250 > gpurun_out/r02_gputests_multi_n2.log 2>&1; tail -30 gpurun_out/r02_gputests_multi_n2.log
echo "NCCL_DEBUG=$NCCL_DEBUG"

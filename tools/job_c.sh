# round 2, job c: the whole GPU suite on HEAD, then the ncu pass (launch lists + full captures of every dominant kernel)
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02_gputests_head.log
bash tools/ncu_job.sh a > gpurun_out/ncu_job.log 2>&1
tail -25 gpurun_out/ncu_job.log

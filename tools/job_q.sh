timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python -m pytest -q -x --timeout 200 -m gpu tests 2>&1 | tail -6 | tee gpurun_out/r02_gputests_final2.log

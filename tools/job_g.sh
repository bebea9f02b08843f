# round 2, job g: what the small-query kernels really cost (ncu durations; the CUDA-event pairs of the profile mode are as long as the kernels)
B="python bench.py --no-e2e --no-cpu --no-sharded --no-poolfirst --no-parity"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_cfg4i_g.csv $B --workload cfg4i --steps 3 --warmup 2 > gpurun_out/ncu_l_cfg4i.log 2>&1
python tools/launch_share.py gpurun_out/r02_launches_cfg4i_g.csv | head -12
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gemv8 -s 4 -c 2 -f -o gpurun_out/r02_cfg4i_gemv_g $B --workload cfg4i --steps 2 --warmup 2 > gpurun_out/ncu_cfg4i_gemv.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_cfg4i_gemv_g.ncu-rep gpurun_out/r02_cfg4i_gemv_g_ncu_summary.json --traffic-key cfg4i --traffic-out gpurun_out/roofline_traffic_g.json

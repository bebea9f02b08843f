set -x
B="python bench.py --no-e2e --no-cpu"
timeout 300 $B --steps 2 --warmup 1 > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_v6.csv $B --steps 2 --warmup 1 > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_poolacc -s 1 -c 1 -f -o gpurun_out/r01_poolacc_v6 $B --steps 1 --warmup 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pa_normalize_scatter -s 1 -c 1 -f -o gpurun_out/r01_normscatter_v6 $B --steps 1 --warmup 1 > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_exact_q30 -s 1 -c 1 -f -o gpurun_out/r01_exact_v6 $B --steps 1 --warmup 1 > gpurun_out/ncu_c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_poolacc -s 1 -c 1 -f -o gpurun_out/r01_cfg5_poolacc_v6 $B --workload cfg5 --steps 1 --warmup 1 > gpurun_out/ncu_d.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches_cfg5_v6.csv $B --workload cfg5 --steps 2 --warmup 1 > gpurun_out/ncu_e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_poolgemm -s 1 -c 1 -f -o gpurun_out/r01_cfg4_poolgemm_v6 $B --workload cfg4 --steps 1 --warmup 1 > gpurun_out/ncu_f.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches_cfg4_v6.csv $B --workload cfg4 --steps 2 --warmup 1 > gpurun_out/ncu_g.log 2>&1
ls -la gpurun_out/ | tail -20
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log gpurun_out/ncu_d.log gpurun_out/ncu_f.log

# round 2, job e: the rewritten small-query path -- GPU suite, then cfg4i / cfg2 / cfg4ii bench lines, launch list + full capture of k_gemv8
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02_gputests_e.log
B="python bench.py --no-cpu --no-sharded"
for w in cfg4i cfg2 cfg4ii; do
  timeout 300 $B --workload $w --steps 20 --warmup 5 > gpurun_out/r02_bench_${w}_n1_e.json 2> gpurun_out/r02_bench_${w}_e.err || tail -5 gpurun_out/r02_bench_${w}_e.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_cfg*_n1_e.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'roof', round(d['roofline']['frac'],3), 'avg', round(d['roofline']['avg_launch_ms'],4), 'par', (d['parity_sample'] or {}).get('status'), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v>0}, 'e2e', '%.3g'%d['e2e']['value'] if d['e2e'] else None, d['clocks']['sm_mhz'], 'launches', d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
R=r02; V=e
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_cfg4i_$V.csv python bench.py --no-e2e --no-cpu --no-sharded --workload cfg4i --steps 2 --warmup 1 > gpurun_out/ncu_l_cfg4i.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemv8 -s 2 -c 2 -f -o gpurun_out/${R}_cfg4i_gemv_$V python bench.py --no-e2e --no-cpu --no-sharded --workload cfg4i --steps 1 --warmup 1 > gpurun_out/ncu_cfg4i_gemv.log 2>&1
python tools/ncu_summary.py gpurun_out/${R}_cfg4i_gemv_$V.ncu-rep gpurun_out/${R}_cfg4i_gemv_${V}_ncu_summary.json --traffic-key cfg4i --traffic-out gpurun_out/roofline_traffic_e.json

# round 2, job h: rewritten stream kernel + tail, pool-first on k_poolacc, cta2 tests; short timeouts everywhere
set -o pipefail
T="timeout 420 python -m pytest -q -x --timeout 100 -m gpu"
$T tests/test_gpu_small.py tests/test_gpu_certificate.py tests/test_gpu_cli.py 2>&1 | tail -12 | tee gpurun_out/r02_gputests_h1.log || { echo "SMALL PATH TESTS FAILED"; exit 1; }
$T tests/test_gpu_poolfirst.py tests/test_gpu_cta2.py 2>&1 | tail -12 | tee gpurun_out/r02_gputests_h2.log
timeout 900 python -m pytest -q -x --timeout 150 -m gpu tests 2>&1 | tail -12 | tee gpurun_out/r02_gputests_h.log
B="python bench.py --no-cpu --no-sharded --no-poolfirst"
for w in cfg4i cfg2; do
  timeout 150 $B --workload $w --steps 20 --warmup 5 > gpurun_out/r02_bench_${w}_n1_h.json 2> gpurun_out/r02_bench_${w}_h.err || tail -5 gpurun_out/r02_bench_${w}_h.err
done
timeout 200 $B --workload cfg3 --no-e2e --stage-a poolfirst --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_poolfirst_n1_h.json 2> gpurun_out/r02_bench_cfg3_pf_h.err || tail -5 gpurun_out/r02_bench_cfg3_pf_h.err
timeout 200 $B --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg5_n1_h.json 2> gpurun_out/r02_bench_cfg5_h.err || tail -5 gpurun_out/r02_bench_cfg5_h.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_cfg*_n1_h.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'roof', round(d['roofline']['frac'],3), 'avg', round(d['roofline']['avg_launch_ms'],4), 'par', (d.get('parity_sample') or {}).get('status'), {k:round(v,4) for k,v in d.get('kernel_ms_per_step',{}).items() if v>0}, 'e2e', '%.3g'%d['e2e']['value'] if d['e2e'] else None, d['clocks']['sm_mhz'], 'launches', d['gpu_launches'], d['config'].get('path'))
    except Exception as e:
        print(f, 'ERR', e)
PY
BB="python bench.py --no-e2e --no-cpu --no-sharded --no-poolfirst --no-parity"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_cfg4i_h.csv $BB --workload cfg4i --steps 3 --warmup 2 > gpurun_out/ncu_l_cfg4i.log 2>&1
python tools/launch_share.py gpurun_out/r02_launches_cfg4i_h.csv | head -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemv8 -s 4 -c 2 -f -o gpurun_out/r02_cfg4i_gemv_h $BB --workload cfg4i --steps 2 --warmup 2 > gpurun_out/ncu_cfg4i_gemv.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_cfg4i_gemv_h.ncu-rep gpurun_out/r02_cfg4i_gemv_h_ncu_summary.json --traffic-key cfg4i --traffic-out gpurun_out/roofline_traffic_h.json

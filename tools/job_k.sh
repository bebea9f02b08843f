# round 2, job k: fast prep + slim selection in k_gemv8, rolled epilogue in k_poolacc2, write+read L2 flush in bench
set -o pipefail
T="timeout 420 python -m pytest -q -x --timeout 100 -m gpu"
$T tests/test_gpu_small.py tests/test_gpu_certificate.py tests/test_gpu_fuzz.py tests/test_gpu_cta2.py tests/test_gpu_poolfirst.py 2>&1 | tail -8 | tee gpurun_out/r02_gputests_k1.log || { echo "TESTS FAILED"; exit 1; }
B="python bench.py --no-cpu --no-sharded --no-poolfirst"
timeout 150 $B --workload cfg4i --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg4i_n1_k.json 2> gpurun_out/r02_bench_cfg4i_k.err || tail -5 gpurun_out/r02_bench_cfg4i_k.err
timeout 150 $B --workload cfg2 --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg2_n1_k.json 2> gpurun_out/r02_bench_cfg2_k.err || tail -5 gpurun_out/r02_bench_cfg2_k.err
timeout 200 $B --workload cfg3 --no-e2e --stage-a poolfirst --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_poolfirst_n1_k.json 2> gpurun_out/r02_bench_cfg3_pf_k.err || tail -5 gpurun_out/r02_bench_cfg3_pf_k.err
timeout 200 $B --workload cfg3 --no-e2e --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_noe2e_n1_k.json 2> gpurun_out/r02_bench_cfg3_k.err || tail -5 gpurun_out/r02_bench_cfg3_k.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_cfg*_n1_k.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'roof', round(d['roofline']['frac'],3), 'avg', round(d['roofline']['avg_launch_ms'],4), 'probe', d['roofline'].get('read_probe',{}).get('ms'), 'par', (d.get('parity_sample') or {}).get('status'), {k:round(v,4) for k,v in d.get('kernel_ms_per_step',{}).items() if v>0}, d['clocks']['sm_mhz'], 'launches', d['gpu_launches'], d['config'].get('path'))
    except Exception as e:
        print(f, 'ERR', e)
PY
BB="python bench.py --no-e2e --no-cpu --no-sharded --no-poolfirst --no-parity"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_cfg4i_k.csv $BB --workload cfg4i --steps 3 --warmup 2 > gpurun_out/ncu_l_cfg4i.log 2>&1
python tools/launch_share.py gpurun_out/r02_launches_cfg4i_k.csv | head -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemv8 -s 4 -c 2 -f -o gpurun_out/r02_cfg4i_gemv_k $BB --workload cfg4i --steps 2 --warmup 2 > gpurun_out/ncu_cfg4i_gemv.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_cfg4i_gemv_k.ncu-rep gpurun_out/r02_cfg4i_gemv_j_ncu_summary.json --traffic-key cfg4i --traffic-out gpurun_out/roofline_traffic_k.json
timeout 600 python -m pytest -q -x --timeout 150 -m gpu tests 2>&1 | tail -6 | tee gpurun_out/r02_gputests_k.log

# round 2, job f: small-query path (fixed ring ownership), pool-first, k_poolacc2 -- every stage under a short timeout
set -o pipefail
T="timeout 420 python -m pytest -q -x --timeout 100 -m gpu"
$T tests/test_gpu_small.py tests/test_gpu_certificate.py 2>&1 | tail -12 | tee gpurun_out/r02_gputests_f1.log || { echo "SMALL PATH TESTS FAILED"; exit 1; }
$T tests/test_gpu_poolfirst.py 2>&1 | tail -12 | tee gpurun_out/r02_gputests_f2.log
timeout 900 python -m pytest -q -x --timeout 150 -m gpu tests 2>&1 | tail -12 | tee gpurun_out/r02_gputests_f.log
B="python bench.py --no-cpu --no-sharded --no-poolfirst"
for w in cfg4i cfg2 cfg4ii; do
  timeout 150 $B --workload $w --steps 20 --warmup 5 > gpurun_out/r02_bench_${w}_n1_f.json 2> gpurun_out/r02_bench_${w}_f.err || tail -5 gpurun_out/r02_bench_${w}_f.err
done
timeout 200 $B --workload cfg3 --no-e2e --stage-a poolfirst --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_poolfirst_n1_f.json 2> gpurun_out/r02_bench_cfg3_pf_f.err || tail -5 gpurun_out/r02_bench_cfg3_pf_f.err
timeout 150 $B --workload cfg3 --no-e2e --cta-group 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_cta2_n1_f.json 2> gpurun_out/r02_bench_cfg3_cta2_f.err || tail -5 gpurun_out/r02_bench_cfg3_cta2_f.err
timeout 200 $B --workload cfg3 --no-e2e --cta-group 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_cta1_n1_f.json 2> gpurun_out/r02_bench_cfg3_cta1_f.err || tail -5 gpurun_out/r02_bench_cfg3_cta1_f.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_cfg*_n1_f.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'roof', round(d['roofline']['frac'],3), 'avg', round(d['roofline']['avg_launch_ms'],4), 'par', (d['parity_sample'] or {}).get('status'), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v>0}, 'e2e', '%.3g'%d['e2e']['value'] if d['e2e'] else None, d['clocks']['sm_mhz'], 'launches', d['gpu_launches'], d['config']['path'])
    except Exception as e:
        print(f, 'ERR', e)
PY

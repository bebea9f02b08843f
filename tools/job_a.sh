# A/B job of round 2: tests, then cfg4 / cfg4i / cfg4ii with and without the running k-th best pruning, cfg3 max pooling
B="python bench.py --no-cpu --no-sharded"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for w in cfg4 cfg4i cfg4ii; do
  for k in 0 1; do
    timeout 300 $B --workload $w --kth $k --steps 10 --warmup 3 > gpurun_out/r02_${w}_kth${k}.json 2> gpurun_out/r02_${w}_kth${k}.err || tail -5 gpurun_out/r02_${w}_kth${k}.err
  done
done
timeout 300 $B --workload cfg3 --pool max --steps 3 --warmup 3 > gpurun_out/r02_cfg3_max_b.json 2> gpurun_out/r02_cfg3_max_b.err || tail -5 gpurun_out/r02_cfg3_max_b.err
timeout 300 $B --workload cfg4 --pool max --steps 5 --warmup 3 > gpurun_out/r02_cfg4_max_b.json 2> gpurun_out/r02_cfg4_max_b.err || tail -5 gpurun_out/r02_cfg4_max_b.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_cfg*_kth*.json')+glob.glob('gpurun_out/r02_cfg*_max_b.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'roof', round(d['roofline']['frac'],3), 'avg', round(d['roofline']['avg_launch_ms'],4), 'par', (d['parity_sample'] or {}).get('status'), 'retry', d['config']['certificate_retry_groups'], 'fb', d['config']['certificate_fallback_groups'], {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v>0}, 'e2e', '%.3g'%d['e2e']['value'] if d['e2e'] else None, d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY

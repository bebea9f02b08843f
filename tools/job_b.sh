B="python bench.py --no-cpu --no-sharded --no-e2e"
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or fuzz or certificate" 2>&1 | tail -3
for k in 0 1; do
  timeout 300 $B --workload cfg4i --kth $k --steps 20 --warmup 5 > gpurun_out/r02_cfg4i_kth${k}_b.json 2> gpurun_out/r02_cfg4i_kth${k}_b.err || tail -5 gpurun_out/r02_cfg4i_kth${k}_b.err
done
timeout 300 $B --workload cfg3 --pool max --steps 3 --warmup 3 > gpurun_out/r02_cfg3_max_nc256.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
SDK_PG_NC=128 timeout 300 $B --workload cfg3 --pool max --steps 3 --warmup 3 > gpurun_out/r02_cfg3_max_nc128.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
SDK_PG_NC=128 timeout 300 $B --workload cfg3 --pool mean --acc 0 --steps 3 --warmup 3 > gpurun_out/r02_cfg3_mean_generic_nc128.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
timeout 300 $B --workload cfg3 --pool mean --acc 0 --steps 3 --warmup 3 > gpurun_out/r02_cfg3_mean_generic_nc256.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_cfg4i_kth*_b.json')+glob.glob('gpurun_out/r02_cfg3_m*_nc*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'roof', round(d['roofline']['frac'],3), 'avg', round(d['roofline']['avg_launch_ms'],4), 'par', (d['parity_sample'] or {}).get('status'), {k:round(v,4) for k,v in d['kernel_ms_per_step'].items() if v>0}, d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY

# A/B: CTA pairs for the generic kernel (config 4 mean pooling at D = 512, config 3 max pooling at D = 192), alternating runs on one box
B="python bench.py --no-cpu --no-sharded --no-poolfirst --no-e2e"
for rep in 1 2; do
  for cg in 1 2; do
    timeout 200 $B --workload cfg4 --cta-group $cg --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg4_cg${cg}_${rep}_v.json 2> gpurun_out/err_v.txt || tail -3 gpurun_out/err_v.txt
    timeout 200 $B --workload cfg3 --pool max --cta-group $cg --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg3max_cg${cg}_${rep}_v.json 2> gpurun_out/err_v.txt || tail -3 gpurun_out/err_v.txt
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_cfg*_cg*_v.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],3), 'par', (d.get('parity_sample') or {}).get('status'), d['clocks']['sm_mhz'], d['config']['path'])
    except Exception as e:
        print(f,'ERR',e)
PY

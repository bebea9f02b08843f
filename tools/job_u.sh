# round 2, final validation of the shipped build: smoke, the whole GPU suite, the driver's default command, the reference arm
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python -m pytest -q -x --timeout 200 -m gpu tests 2>&1 | tail -5 | tee gpurun_out/r02_gputests_final3.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n1_final3.json 2> gpurun_out/r02_bench_cfg3_final3.err; echo "cfg3 rc=$?"; tail -2 gpurun_out/r02_bench_cfg3_final3.err
B="python bench.py --no-sharded --no-poolfirst --no-cpu"
timeout 200 $B --workload cfg4i --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg4i_n1_final3.json 2>/dev/null
timeout 200 $B --workload cfg5 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg5_n1_final3.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*_n1_final3.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'val %.3g'%d['value'], 'roof', r.get('frac') and round(r['frac'],3), 'traffic', r.get('traffic'), 'e2e', d.get('e2e') and '%.3g'%d['e2e']['value'], 'e2e_f32', d.get('e2e_f32') and '%.3g'%d['e2e_f32']['value'], 'par', (d.get('parity_sample') or {}).get('status'), (d.get('parity_sample_e2e') or {}).get('status'), 'cpu', d.get('cpu_baseline') and '%.3g'%d['cpu_baseline']['value'], (d.get('clocks') or {}).get('sm_mhz'))
        if 'sharded' in d: print('   sharded', d['sharded']['ms_per_step'], d['sharded']['roofline']['frac'], d['sharded']['parity_sample'])
        if 'pool_first' in d: print('   pool_first', d['pool_first']['time_to_solution_ms'], (d['pool_first']['parity_sample'] or {}).get('status'))
    except Exception as e:
        print(f, 'ERR', e)
PY

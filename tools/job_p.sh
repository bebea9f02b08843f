# round 2, job p: the final build -- whole GPU suite, every bench line, then the ncu pass
timeout 900 python -m pytest -q -x --timeout 150 -m gpu tests 2>&1 | tail -6 | tee gpurun_out/r02_gputests_final.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n1_final.json 2> gpurun_out/r02_bench_cfg3_final.err; echo "cfg3 rc=$?"; tail -2 gpurun_out/r02_bench_cfg3_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n1_final.json 2> gpurun_out/r02_bench_reference_final.err; echo "reference rc=$?"
B="python bench.py --no-sharded --no-poolfirst"
for w in cfg1 cfg2 cfg4 cfg4i cfg4ii cfg5; do
  timeout 200 $B --workload $w --steps 10 --warmup 3 > gpurun_out/r02_bench_${w}_n1_final.json 2> gpurun_out/r02_bench_${w}_final.err || tail -3 gpurun_out/r02_bench_${w}_final.err
done
timeout 300 $B --workload cfg3 --pool max --no-cpu --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg3_max_n1_final.json 2> gpurun_out/r02_bench_cfg3_max_final.err || tail -3 gpurun_out/r02_bench_cfg3_max_final.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*_n1_final.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'val %.3g'%d['value'], 'roof', r.get('frac') and round(r['frac'],3), 'e2e', d.get('e2e') and '%.3g'%d['e2e']['value'], 'e2e_f32', d.get('e2e_f32') and '%.3g'%d['e2e_f32']['value'], 'par', (d.get('parity_sample') or {}).get('status'), 'cpu', d.get('cpu_baseline') and '%.3g'%d['cpu_baseline']['value'], (d.get('clocks') or {}).get('sm_mhz'))
        if 'sharded' in d: print('   sharded', d['sharded']['ms_per_step'], d['sharded']['roofline']['frac'], d['sharded']['parity_sample'])
        if 'pool_first' in d: print('   pool_first', d['pool_first']['time_to_solution_ms'], (d['pool_first']['parity_sample'] or {}).get('status'))
    except Exception as e:
        print(f, 'ERR', e)
PY
R=r02 bash tools/ncu_job.sh z > gpurun_out/ncu_job_z.log 2>&1
tail -12 gpurun_out/ncu_job_z.log

# the shipped build on one GPU: smoke, the whole GPU suite, the driver's default bench command, the small-query bench line
rm -f speaker_diarization_toolkit_b200/libsdk_b200_trace*.so
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/r02_smoke_final4.log
timeout 900 python -m pytest -q -x --timeout 200 -m gpu tests 2>&1 | tail -6 | tee gpurun_out/r02_gputests_final4.log
timeout 400 python bench.py > gpurun_out/r02_bench_cfg3_n1_final4.json 2> gpurun_out/r02_bench_cfg3_final4.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_cfg3_final4.err
timeout 150 python bench.py --no-cpu --no-sharded --no-poolfirst --workload cfg4i --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg4i_n1_final4.json 2> gpurun_out/r02_bench_cfg4i_final4.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_cfg3_n1_final4.json','gpurun_out/r02_bench_cfg4i_n1_final4.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'value %.3g'%d['value'], 'roof', round(d['roofline']['frac'],3), 'par', (d.get('parity_sample') or {}).get('status'), 'e2e', d['e2e'] and '%.3g'%d['e2e']['value'], 'launches', d['gpu_launches'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
        s=d.get('sharded')
        if s: print('  sharded ms', s['ms_per_step'], 'frac', s['roofline']['frac'], 'par', s.get('parity_sample'))
        pf=d.get('pool_first')
        if pf: print('  pool_first ms', pf['time_to_solution_ms'], (pf['parity_sample'] or {}).get('status'))
        print('  cpu', d.get('cpu_baseline'))
    except Exception as e:
        print(f, 'ERR', e)
PY

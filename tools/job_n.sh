# 2 GPUs: the product CLI on 2 GPUs (full error output), then the driver's bench command with both e2e legs
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 250 > gpurun_out/r02_gputests_multi_n2.log 2>&1; tail -40 gpurun_out/r02_gputests_multi_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 420 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n2_n.json 2> gpurun_out/r02_bench_cfg3_n2_n.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_cfg3_n2_n.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_cfg3_n2_n.json').read().strip().splitlines()[-1])
    print('cfg3 ms', d['ms_per_step'], 'value %.3g'%d['value'], 'e2e', d['e2e'] and ('%.3g'%d['e2e']['value'], d['e2e'].get('h2d_gbs_per_rank'), d['e2e'].get('numa_node_of_gpu')), 'e2e_f32', d.get('e2e_f32') and '%.3g'%d['e2e_f32']['value'], 'par', (d['parity_sample'] or {}).get('status'), (d.get('parity_sample_e2e') or {}).get('status'))
    s=d.get('sharded')
    if s: print('sharded ms', s['ms_per_step'], 'value %.3g'%s['value'], 'frac', s['roofline']['frac'], 'ag us', s.get('allgather_merge_us'), 'par', s.get('parity_sample'), 'e2e', s['e2e'] and '%.3g'%s['e2e']['value'])
    pf=d.get('pool_first')
    if pf: print('pool_first ms', pf['time_to_solution_ms'], 'par', (pf['parity_sample'] or {}).get('status'), pf['roofline']['frac'])
except Exception as e:
    print('ERR', e)
PY

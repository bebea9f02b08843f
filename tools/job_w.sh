# 8 GPUs, the driver's command on the shipped build (e2e legs with NUMA placement by sysfs / H2D probe)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
lscpu | grep -i -E "numa|socket|^CPU\(s\)" | head -8
nvidia-smi topo -m 2>/dev/null | head -14
timeout 420 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n${N}_w.json 2> gpurun_out/r02_bench_cfg3_n${N}_w.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_cfg3_n${N}_w.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_cfg3_n${N}_w.json').read().strip().splitlines()[-1])
    print('cfg3 ms', d['ms_per_step'], 'value %.3g'%d['value'], 'par', (d['parity_sample'] or {}).get('status'), (d.get('parity_sample_e2e') or {}).get('status'))
    print(' e2e', d['e2e'])
    print(' e2e_f32', d.get('e2e_f32'))
    s=d.get('sharded')
    if s: print('sharded ms', s['ms_per_step'], 'value %.3g'%s['value'], 'frac', s['roofline']['frac'], 'whole', s['roofline'].get('whole_step_frac'), 'ag us', s.get('allgather_merge_us'), 'par', s.get('parity_sample'), 'e2e', s['e2e'] and '%.3g'%s['e2e']['value'])
    pf=d.get('pool_first')
    if pf: print('pool_first ms', pf['time_to_solution_ms'], 'par', (pf['parity_sample'] or {}).get('status'))
except Exception as e:
    print('ERR', e)
PY

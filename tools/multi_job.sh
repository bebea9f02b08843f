# usage: bash tools/multi_job.sh N [TAG]  -- on N GPUs of one box: the sharded parity check, the driver's bench command (default
# workload + row-sharded sub-record + pool-first sub-record), the multi-GPU tests of the product CLI.  Every stage under a timeout.
N=${1:-2}
TAG=${2:-a}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/sharded_check.py > gpurun_out/r02_sharded_check_n${N}.log 2>&1; echo "sharded_check rc=$?"; grep -E "OK|MISMATCH|UNEXPECTED|Error|error" gpurun_out/r02_sharded_check_n${N}.log | tail -40
timeout 420 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n${N}_${TAG}.json 2> gpurun_out/r02_bench_cfg3_n${N}_${TAG}.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_cfg3_n${N}_${TAG}.err
timeout 240 python -m pytest tests -m gpu -x -q --timeout 200 -k "multi" 2>&1 | tail -5 | tee gpurun_out/r02_gputests_multi_n${N}.log
if [ "$3" = "ref" ]; then
  timeout 200 $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n${N}_${TAG}.json 2> gpurun_out/r02_bench_reference_n${N}_${TAG}.err; echo "reference rc=$?"
fi
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_cfg3_n${N}_${TAG}.json').read().strip().splitlines()[-1])
    print('cfg3 ms', d['ms_per_step'], 'value %.3g'%d['value'], 'e2e', d['e2e'] and '%.3g'%d['e2e']['value'], 'e2e_f32', d.get('e2e_f32') and '%.3g'%d['e2e_f32']['value'], 'h2d GB/s/rank', d['e2e'] and d['e2e'].get('h2d_gbs_per_rank'), 'numa', d['e2e'] and d['e2e'].get('numa_node_of_gpu'), 'par', (d['parity_sample'] or {}).get('status'))
    s=d.get('sharded')
    if s: print('sharded ms', s['ms_per_step'], 'value %.3g'%s['value'], 'frac', s['roofline']['frac'], 'whole', s['roofline'].get('whole_step_frac'), 'ag us', s.get('allgather_merge_us'), 'par', s.get('parity_sample'))
    pf=d.get('pool_first')
    if pf: print('pool_first ms', pf['time_to_solution_ms'], 'par', (pf['parity_sample'] or {}).get('status'), pf['roofline']['frac'])
except Exception as e:
    print('ERR', e)
PY

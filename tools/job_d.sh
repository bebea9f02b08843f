# round 2, job d: GPU suite (incl. fp16 inputs, assign-batch --gpus), cfg1 through the CLI, default bench with both e2e legs
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r02_gputests_d.log
timeout 300 python bench.py --workload cfg1 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg1_n1_d.json 2> gpurun_out/r02_bench_cfg1_d.err; echo "cfg1 rc=$?"; tail -3 gpurun_out/r02_bench_cfg1_d.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg3_n1_d.json 2> gpurun_out/r02_bench_cfg3_d.err; echo "cfg3 rc=$?"; tail -3 gpurun_out/r02_bench_cfg3_d.err
python - <<'PY'
import json
for f in ['gpurun_out/r02_bench_cfg1_n1_d.json','gpurun_out/r02_bench_cfg3_n1_d.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', d['ms_per_step'], 'value %.3g'%d['value'], 'e2e', d['e2e'], 'e2e_f32', d.get('e2e_f32'), 'par', d.get('parity_sample'), d.get('parity_sample_e2e'))
        if 'sharded' in d: print(' sharded', d['sharded']['ms_per_step'], d['sharded']['e2e'], d['sharded'].get('e2e_f32'))
    except Exception as e:
        print(f,'ERR',e)
PY

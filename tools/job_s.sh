# group schedule of k_poolacc2: parity, then A/B against the round-robin schedule on one box, then the DRAM traffic
set -o pipefail
timeout 500 python -m pytest -q -x --timeout 150 -m gpu tests/test_gpu_cta2.py tests/test_gpu_poolfirst.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py 2>&1 | tail -5 | tee gpurun_out/r02_gputests_s1.log || { echo TESTS FAILED; exit 1; }
B="python bench.py --no-cpu --no-sharded --no-poolfirst --no-e2e --workload cfg3 --steps 5 --warmup 3"
for rep in 1 2; do
  for sch in 0 1; do
    SDK_PA_SCHED=$sch timeout 200 $B > gpurun_out/r02_bench_cfg3_sched${sch}_${rep}_s.json 2> gpurun_out/err_s.txt || tail -3 gpurun_out/err_s.txt
  done
done
SDK_PA_SCHED=1 timeout 200 $B --stage-a poolfirst > gpurun_out/r02_bench_cfg3_poolfirst_sched1_s.json 2> gpurun_out/err_s.txt || tail -3 gpurun_out/err_s.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_cfg3_*_s.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],3), 'par', (d.get('parity_sample') or {}).get('status'), d['clocks']['sm_mhz'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.3})
    except Exception as e:
        print(f,'ERR',e)
PY
BB="python bench.py --no-e2e --no-cpu --no-sharded --no-poolfirst --no-parity --workload cfg3 --steps 1 --warmup 1"
SDK_PA_SCHED=1 timeout 300 ncu --set full --clock-control none -k regex:k_poolacc -s 1 -c 1 -f -o /tmp/r02_poolacc_sched1 $BB > gpurun_out/ncu_sched1.log 2>&1
python tools/ncu_summary.py /tmp/r02_poolacc_sched1.ncu-rep gpurun_out/r02_poolacc_sched1_ncu_summary.json --traffic-key cfg3 --traffic-out gpurun_out/roofline_traffic_s.json

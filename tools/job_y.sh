# round 2, job y: small-query path, second step (B fragments written by the normalising warps, labels per warp) + the whole GPU suite
set -o pipefail
S=${1:-y}
T="timeout 420 python -m pytest -q -x --timeout 100 -m gpu"
$T tests/test_gpu_small.py tests/test_gpu_certificate.py tests/test_gpu_f16.py tests/test_gpu_fuzz.py 2>&1 | tail -12 | tee gpurun_out/r02_gputests_${S}1.log || { echo "SMALL PATH TESTS FAILED"; exit 1; }
timeout 120 python tools/gv_trace.py 2>&1 | tee gpurun_out/gv_trace_${S}.log
B="python bench.py --no-cpu --no-sharded --no-poolfirst"
timeout 150 $B --workload cfg4i --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg4i_n1_${S}.json 2> gpurun_out/r02_bench_cfg4i_${S}.err || tail -5 gpurun_out/r02_bench_cfg4i_${S}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_cfg4i_n1_${S}.json').read().strip().splitlines()[-1])
print('cfg4i ms', round(d['ms_per_step'],4), 'roof', round(d['roofline']['frac'],3), 'avg', round(d['roofline']['avg_launch_ms'],4), 'par', (d.get('parity_sample') or {}).get('status'), {k:round(v,4) for k,v in d.get('kernel_ms_per_step',{}).items() if v>0}, 'e2e', '%.3g'%d['e2e']['value'] if d['e2e'] else None, d['clocks']['sm_mhz'])
PY
BB="python bench.py --no-e2e --no-cpu --no-sharded --no-poolfirst --no-parity"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_cfg4i_${S}.csv $BB --workload cfg4i --steps 3 --warmup 2 > gpurun_out/ncu_l_cfg4i.log 2>&1
python tools/launch_share.py gpurun_out/r02_launches_cfg4i_${S}.csv | head -8
timeout 900 python -m pytest -q -x --timeout 150 -m gpu tests 2>&1 | tail -6 | tee gpurun_out/r02_gputests_${S}.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemv8 -s 4 -c 2 -f -o gpurun_out/r02_cfg4i_gemv_${S} $BB --workload cfg4i --steps 2 --warmup 2 > gpurun_out/ncu_cfg4i_gemv.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_cfg4i_gemv_${S}.ncu-rep gpurun_out/r02_cfg4i_gemv_${S}_ncu_summary.json --traffic-key cfg4i --traffic-out gpurun_out/roofline_traffic_${S}.json
timeout 150 $B --workload cfg2 --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg2_n1_${S}.json 2> gpurun_out/r02_bench_cfg2_${S}.err || tail -5 gpurun_out/r02_bench_cfg2_${S}.err

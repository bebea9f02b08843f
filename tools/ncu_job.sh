# Profiling pass (run on the GPU box: `gpurun -- bash tools/ncu_job.sh a`; files are named ${R}_..., R = round prefix, default r02).  Every ncu command runs only after
# the same bench command has exited 0 without ncu.  Reports land in gpurun_out/; summarise with tools/ncu_summary.py.
V=${1:-a}
R=${R:-r02}
B="python bench.py --no-e2e --no-cpu --no-sharded --no-poolfirst --no-parity"
set -x
for w in cfg3 cfg4 cfg5 cfg4i cfg4ii; do
  timeout 200 $B --workload $w --steps 2 --warmup 1 > gpurun_out/ncu_plain_$w.json 2> gpurun_out/ncu_plain_$w.err || exit 1
done
timeout 200 $B --workload cfg3 --stage-a poolfirst --steps 2 --warmup 1 > gpurun_out/ncu_plain_cfg3pf.json 2> gpurun_out/ncu_plain_cfg3pf.err || exit 1
# launch lists (share of the step per kernel)
for w in cfg3 cfg4 cfg5 cfg4i; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_${w}_$V.csv $B --workload $w --steps 2 --warmup 1 > gpurun_out/ncu_l_$w.log 2>&1
done
# full captures of the dominant kernels (one launch each, after the warm-up launches)
# (gpurun copies at most 64 MiB back: the reports are summarised here and only the first one is kept)
cap() {  # workload kernel-regex name [traffic-key]
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o /tmp/${R}_$3_$V $B --workload $1 $X --steps 1 --warmup 1 > gpurun_out/ncu_$3.log 2>&1
  python tools/ncu_summary.py /tmp/${R}_$3_$V.ncu-rep gpurun_out/${R}_$3_${V}_ncu_summary.json ${4:+--traffic-key $4} --traffic-out gpurun_out/roofline_traffic.json
}
cap cfg3 k_poolacc poolacc cfg3
cp /tmp/${R}_poolacc_$V.ncu-rep gpurun_out/
cap cfg3 k_pa_normalize_scatter normscatter
cap cfg3 k_exact_q30 exact
cap cfg3 k_pg_merge merge
cap cfg5 k_poolacc cfg5_poolacc cfg5
cap cfg4 "^k_poolgemm$" cfg4_poolgemm cfg4
cap cfg4i "^k_gemv8$" cfg4i_gemv cfg4i
cap cfg4i k_gemv8_tail cfg4i_tail
cap cfg4ii "^k_poolgemm$" cfg4ii_poolgemm cfg4ii
cap cfg4 k_pg_merge cfg4_merge
cp /tmp/${R}_cfg4i_gemv_$V.ncu-rep gpurun_out/
X="--pool max" cap cfg3 "^k_poolgemm$" cfg3_max_poolgemm cfg3_max
X="--stage-a poolfirst" cap cfg3 k_normalize_centroid cfg3_poolfirst_k1 cfg3_poolfirst
X="--stage-a poolfirst" cap cfg3 k_poolacc2 cfg3_poolfirst_gemm
ls -la gpurun_out/ | tail -30

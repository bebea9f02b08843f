# experiment: where do k_poolacc2's extra DRAM reads come from?  full captures of the config-3 launch with single CTAs / CTA
# pairs on all SMs and on a grid that is a multiple of the units per block (no block straddles two rounds of the static schedule)
B="python bench.py --no-e2e --no-cpu --no-sharded --no-poolfirst --no-parity --workload cfg3 --steps 1 --warmup 1"
cap() {  # tag cta_group [grid]
  SDK_PA_GRID=$3 timeout 300 ncu --set full --clock-control none -k regex:k_poolacc -s 1 -c 1 -f -o /tmp/r02_poolacc_$1 $B --cta-group $2 > gpurun_out/ncu_exp_$1.log 2>&1
  python tools/ncu_summary.py /tmp/r02_poolacc_$1.ncu-rep gpurun_out/r02_poolacc_exp_$1_ncu_summary.json
}
cap cta1_148 1 148
cap cta2_148 2 148
cap cta2_120 2 120
cap cta1_120 1 120
cap cta2_80 2 80

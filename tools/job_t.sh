# trace A/B of the stream kernel: control / no prefetch for the second pass / + L2::256B loads / L2::256B loads with prefetch
for v in "" 2 3 4; do
  timeout 120 python tools/gv_trace.py --lib libsdk_b200_trace$v.so 2>&1 | head -19 | tee gpurun_out/gv_trace_g$v.log
done

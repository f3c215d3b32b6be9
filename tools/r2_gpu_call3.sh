mkdir -p gpurun_out
for f in 0; do
( DBG_FLAGS=$f timeout 300 python tools/gpu_decode_trace.py ) > gpurun_out/r2_decode_trace_f$f.log 2>&1
echo "== DBG_FLAGS=$f"; cat gpurun_out/r2_decode_trace_f$f.log | head -60
done

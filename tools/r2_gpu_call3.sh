mkdir -p gpurun_out
( timeout 300 python tools/gpu_decode_trace.py ) > gpurun_out/r2_decode_trace.log 2>&1
cat gpurun_out/r2_decode_trace.log | tail -60

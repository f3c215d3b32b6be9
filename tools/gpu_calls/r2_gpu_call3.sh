mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_generate.py -q -x -k "default or loop" 2>&1 | tail -5 )
( DBG_FLAGS=0 timeout 300 python tools/gpu_decode_trace.py ) > gpurun_out/r2_decode_trace_f0.log 2>&1
cat gpurun_out/r2_decode_trace_f0.log | head -30

mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_generate.py -q -x -k "default_model_teacher_forced" 2>&1 | tail -6 )
( timeout 300 python tools/gpu_decode_trace64.py ) 2>&1 | tail -32

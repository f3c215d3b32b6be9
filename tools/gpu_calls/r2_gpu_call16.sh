#!/bin/bash
echo "== fwd3 (two Q tiles per CTA)"; timeout 300 python tools/gpu_attn_check.py 2>&1 | grep -v Warn | tail -13
echo "== trace"; PIANOBART_B200_LIB=pianobart_b200/libpianobart_b200_trace.so timeout 200 python tools/gpu_attn_trace2.py 2>&1 | tail -30
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "attention or flash or default_model" 2>&1 | tail -3

#!/bin/bash
export PIANOBART_B200_ATTN_FWD=4
echo "== fwd4 (two Q tiles, 16 softmax warps)"; timeout 300 python tools/gpu_attn_check.py 2>&1 | grep -v "Warn\|return Var" | tail -13
echo "== trace"; PIANOBART_B200_LIB=pianobart_b200/libpianobart_b200_trace.so timeout 200 python tools/gpu_attn_trace2.py 2>&1 | tail -30
echo "== fwd3"; PIANOBART_B200_ATTN_FWD=3 timeout 300 python tools/gpu_attn_check.py 2>&1 | grep "bench fwd"

#!/bin/bash
# two-CTA-per-SM attention forward: correctness + timing A/B, then the attention parity tests
echo "== fwd2 (two CTAs per SM)"; timeout 300 python tools/gpu_attn_check.py 2>&1 | tail -14
echo "== look-ahead kernel"; PIANOBART_B200_ATTN_FWD2=0 timeout 300 python tools/gpu_attn_check.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "attention or flash or default_model" 2>&1 | tail -5

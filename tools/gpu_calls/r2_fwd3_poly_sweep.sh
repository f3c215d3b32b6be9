for m in 0x1111 0x2492 0x5252; do
  echo "== mask $m"
  PIANOBART_B200_LIB=pianobart_b200/libpianobart_b200_trace_$m.so timeout 200 python tools/gpu_attn_trace2.py 2>&1 | grep -A6 "softmax warp 1" | tail -3
  PIANOBART_B200_LIB=pianobart_b200/libpianobart_b200_trace_$m.so timeout 200 python tools/gpu_attn_check.py 2>&1 | grep "bench fwd"
done

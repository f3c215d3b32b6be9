#!/bin/bash
# front-end products as single GEMMs + packed / polynomial softmax in the attention backward kernels
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4 ) | tee gpurun_out/r2_pytest_parity_c65.log
for v in base "" p0 p50; do
  if [ -z "$v" ]; then lib=pianobart_b200/libpianobart_b200.so; else lib=pianobart_b200/libpianobart_b200_$v.so; fi
  echo "== $lib"
  ( PIANOBART_B200_LIB=$PWD/$lib timeout 300 python tools/gpu_attn_check.py 2>&1 | grep -v Warn | tail -14 )
done 2>&1 | tee gpurun_out/r2_attn_bwd_poly.log
( timeout 600 python tools/gpu_gemm_prof.py 2>&1 | grep -i "front\|total" ) | tee gpurun_out/r2_gemm_front_sites.txt
( timeout 1200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2_bench_c65.json; cut -c1-330 gpurun_out/r2_bench_c65.json
( PIANOBART_B200_LIB=$PWD/pianobart_b200/libpianobart_b200_base.so timeout 1200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2_bench_c65_base.json; cut -c1-330 gpurun_out/r2_bench_c65_base.json

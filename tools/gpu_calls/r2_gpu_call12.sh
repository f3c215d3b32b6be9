#!/bin/bash
# fused heads + CE: parity, timing, whole-step check
timeout 600 python -m pytest tests/test_gpu_heads.py -x -q -k "fused" 2>&1 | tail -15 > gpurun_out/r2_pytest_fusedce.log
cat gpurun_out/r2_pytest_fusedce.log
timeout 300 python tools/gpu_heads_ce_bench.py 2>&1 | tail -5 | tee gpurun_out/r2_heads_ce_bench.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2_pytest_parity_fusedce.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-decode --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_fusedce.json | cut -c1-400
PIANOBART_B200_FUSED_CE=0 timeout 600 python bench.py --steps 40 --warmup 5 --no-decode --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_unfusedce.json | cut -c1-400

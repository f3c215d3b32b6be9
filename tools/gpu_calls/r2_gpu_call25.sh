#!/bin/bash
timeout 300 python tools/gpu_attn_check.py 2>&1 | grep -v "Warn\|return Var" | tail -13
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 40 --warmup 5 --no-decode --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_fwd3.json | cut -c1-300

#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_finetune.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 40 --warmup 5 --no-decode --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_side.json | cut -c1-200
PIANOBART_B200_SIDE_COLSUM=0 timeout 600 python bench.py --steps 40 --warmup 5 --no-decode --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_bench_noside.json | cut -c1-200

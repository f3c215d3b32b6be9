#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_final2.csv python bench.py --steps 1 --warmup 1 --no-decode --no-cpu-baseline > gpurun_out/r2_bench_under_ncu2.log 2>&1
grep -c "gpu__time_duration" gpurun_out/r2_launches_final2.csv

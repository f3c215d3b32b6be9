#!/bin/bash
# ncu launch list of the bench command (training step only: --no-decode), per-launch durations, no clock control
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-decode > gpurun_out/r2_ncu_bench.log 2>&1; tail -c 200 gpurun_out/r2_ncu_bench.log; wc -l gpurun_out/r2_launches_final.csv

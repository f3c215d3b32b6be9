#!/bin/bash
# full validation of the round-2 build
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) | tee gpurun_out/r2_pytest_full.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) | tee gpurun_out/r2_smoke.log
( timeout 1200 python bench.py 2>&1 | tail -1 ) > gpurun_out/r2_bench_final.json; cut -c1-400 gpurun_out/r2_bench_final.json
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/r2_bench_reference.json; cut -c1-300 gpurun_out/r2_bench_reference.json
for w in genft seqcls tokcls; do ( timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2_bench_$w.json; cut -c1-250 gpurun_out/r2_bench_$w.json; done

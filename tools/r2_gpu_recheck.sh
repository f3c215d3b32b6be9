#!/bin/bash
# re-check after the last host-side changes: full GPU suite, smoke, one default bench line
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) | tee gpurun_out/r2_pytest_full.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 ) | tee gpurun_out/r2_smoke.log
( timeout 1200 python bench.py --no-cpu-baseline --no-decode 2>&1 | tail -1 ) > gpurun_out/r2_bench_recheck.json; cut -c1-260 gpurun_out/r2_bench_recheck.json

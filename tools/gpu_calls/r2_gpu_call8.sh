mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2_pytest_d.log 2>&1
tail -6 gpurun_out/r2_pytest_d.log
( timeout 900 python bench.py --steps 60 --warmup 5 ) > gpurun_out/r2_bench_d.log 2>&1
tail -1 gpurun_out/r2_bench_d.log

mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r2_pytest_c.log 2>&1
tail -15 gpurun_out/r2_pytest_c.log
( timeout 900 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-decode ) > gpurun_out/r2_bench_c.log 2>&1
tail -1 gpurun_out/r2_bench_c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['loss'], d['gpu_launches'], d['clocks'])"

mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm or default_model or fused_step" 2>&1 | tail -8 )
( timeout 900 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-decode ) > gpurun_out/r2_bench_b.log 2>&1
tail -1 gpurun_out/r2_bench_b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['loss'], d['clocks'])"
( PIANOBART_B200_TAIL_SPLIT=0 timeout 900 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-decode ) > gpurun_out/r2_bench_b0.log 2>&1
tail -1 gpurun_out/r2_bench_b0.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['loss'], d['clocks'])"
( timeout 300 python tools/gpu_decode_bench.py 1 ) 2>&1 | tail -1

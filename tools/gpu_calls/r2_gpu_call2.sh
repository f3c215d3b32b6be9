set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_generate.py -q -x 2>&1 | tail -30 ) > gpurun_out/r2_pytest_b.log 2>&1
tail -30 gpurun_out/r2_pytest_b.log
( timeout 300 python tools/gpu_decode_bench.py 1 ) > gpurun_out/r2_decode_b1.log 2>&1
tail -5 gpurun_out/r2_decode_b1.log
( PIANOBART_B200_DECODE_PERSIST=0 timeout 300 python tools/gpu_decode_bench.py 1 ) > gpurun_out/r2_decode_b1_old.log 2>&1
tail -2 gpurun_out/r2_decode_b1_old.log
( timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-decode ) > gpurun_out/r2_bench_a.log 2>&1
tail -2 gpurun_out/r2_bench_a.log
( bash tools/cold_start_diag.sh pianobart_b200/libpianobart_b200_r1.so 30 r1base ) > gpurun_out/r2_cold_start_r1base.log 2>&1
cat gpurun_out/r2_cold_start_r1base.log

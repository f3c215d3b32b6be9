#!/bin/bash
# 2-GPU check of the final build: NCCL data-parallel parity test + the bench line at N = 2 and N = 1 on the same box
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 ) | tee gpurun_out/r2_pytest_2gpu_final.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline --no-decode > gpurun_out/r2_bench_2gpu_final.log 2>&1 ); echo "rc=$?"
grep '"metric"' gpurun_out/r2_bench_2gpu_final.log | tail -1 | cut -c1-300
( timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>&1 | tail -1 ) > gpurun_out/r2_bench_1gpu_same_box.json; cut -c1-200 gpurun_out/r2_bench_1gpu_same_box.json

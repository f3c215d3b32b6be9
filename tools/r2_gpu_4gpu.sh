#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu-baseline --no-decode > gpurun_out/r2_bench_4gpu.log 2>&1 ); echo "rc=$?"
grep '"metric"' gpurun_out/r2_bench_4gpu.log | tail -1 | cut -c1-260

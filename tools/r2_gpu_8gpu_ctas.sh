#!/bin/bash
# 8 GPUs: NCCL CTA cap sweep for the gradient all-reduce (default / 8 / 4 CTAs per collective)
mkdir -p gpurun_out
port=29600
for c in default 8 4; do
  port=$((port+1))
  if [ $c = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  ( timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-decode > gpurun_out/r2_bench_8gpu_ctas_$c.log 2>&1 ); echo "ctas=$c rc=$?"
  grep '"metric"' gpurun_out/r2_bench_8gpu_ctas_$c.log | tail -1 | cut -c1-220
done

#!/bin/bash
# clip-norm partials behind the gradient-final markers / all-reduce buckets: parity + same-box A/B (1 and 2 GPUs)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_finetune.py tests/test_gpu_multi.py -x -q 2>&1 | tail -3 ) | tee gpurun_out/r2_pytest_norm_partials.log
for v in 1 0 1 0; do
  ( PIANOBART_B200_NORM_PARTIALS=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>&1 | tail -1 | cut -c1-180 | sed "s/^/partials=$v /" )
done | tee gpurun_out/r2_norm_partials_ab.log
for v in 1 0; do
  ( PIANOBART_B200_NORM_PARTIALS=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2959$v bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>&1 | grep '"metric"' | tail -1 | cut -c1-180 | sed "s/^/2gpu partials=$v /" )
done | tee -a gpurun_out/r2_norm_partials_ab.log

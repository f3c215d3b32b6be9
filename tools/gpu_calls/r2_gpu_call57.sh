mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline --no-decode > gpurun_out/r2_bench_8gpu_align.log 2>&1 )
echo "rc=$?"
grep -v "Warning\|warn" gpurun_out/r2_bench_8gpu_align.log | tail -40 | cut -c1-400

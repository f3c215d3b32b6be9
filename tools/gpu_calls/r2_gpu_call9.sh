mkdir -p gpurun_out
nvidia-smi -L
( timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5 ) > gpurun_out/r2_pytest_2gpu.log 2>&1
cat gpurun_out/r2_pytest_2gpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2_bench_2gpu.log 2>&1
tail -1 gpurun_out/r2_bench_2gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'])"
( NCCL_MAX_CTAS=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline ) > gpurun_out/r2_bench_2gpu_ctas4.log 2>&1
tail -1 gpurun_out/r2_bench_2gpu_ctas4.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('max_ctas=4', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'])"

mkdir -p gpurun_out
nvidia-smi -L | wc -l
for cfg in "PIANOBART_B200_COMM_ALIGN=1" "PIANOBART_B200_COMM_ALIGN=0"; do
  echo "== 8 GPUs $cfg"
  ( env $cfg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>&1 | tail -1 > gpurun_out/r2_bench_8gpu_$cfg.json; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_8gpu_$cfg.json').read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])" )
done
echo "== 1 GPU"
( timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])" )

for rep in 1 2; do
for mb in 48 32 24; do
  echo "== bucket $mb MB"
  ( PIANOBART_B200_BUCKET_MB=$mb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$rep bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])" )
done
done

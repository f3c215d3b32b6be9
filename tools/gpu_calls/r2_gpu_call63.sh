export PIANOBART_B200_STEP_GRAPH_DP=1
( timeout 300 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 )
for rep in 1 2; do
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$rep bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>gpurun_out/r2_dp_graph_err_$rep.log | grep "^{" | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('graph', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'], d['gpu_launches'])" ); tail -3 gpurun_out/r2_dp_graph_err_$rep.log | cut -c1-300
done
export PIANOBART_B200_STEP_GRAPH_DP=0
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29639 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>/dev/null | grep "^{" | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('eager', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])" )
( timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline --no-decode 2>/dev/null | grep "^{" | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1gpu', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])" )

#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_finetune.py -x -q 2>&1 | tail -3
for rep in 1 2; do
for cfg in "PIANOBART_B200_STEP_GRAPH=1" "PIANOBART_B200_STEP_GRAPH=0"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --steps 40 --warmup 5 --no-decode --no-cpu-baseline 2>&1 | grep "^{\|capture failed" | tail -2 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])
    else: print(l.strip()[:200])"
done
done

#!/bin/bash
for rep in 1 2; do
for cfg in "" "PIANOBART_B200_SIDE_PREP=0" "PIANOBART_B200_SIDE_COLSUM=0"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --steps 40 --warmup 5 --no-decode --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
done
done

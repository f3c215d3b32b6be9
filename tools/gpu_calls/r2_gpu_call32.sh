#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_finetune.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for rep in 1 2; do
  timeout 600 python bench.py --steps 60 --warmup 5 --no-decode --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"
done

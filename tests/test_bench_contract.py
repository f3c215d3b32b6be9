"""bench.py prints ONE JSON line with the keys the driver reads.  CPU: the reference arm (`--impl reference`, the reference's
own CPU implementation of the path where its tree exists, else the oracle port).  GPU: the product arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, timeout=timeout,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    j = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0'], 600)
    assert j['impl'] == 'reference' and j['metric'] == 'pretrain_octuple_tokens_per_s' and j['unit'] == 'tokens/s'
    assert j['higher_is_better'] is True and j['n_gpus'] == 1 and j['value'] > 0 and j['ms_per_step'] > 0
    cb = j['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['sample'] and cb['value'] == j['value']
    e = j['e2e']
    assert e['value'] == j['value'] and e['unit'] == j['unit'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


@pytest.mark.gpu
def test_product_arm_line():
    j = _run(['--steps', '20', '--warmup', '3', '--no-cpu-baseline', '--no-decode'], 900)
    assert 'impl' not in j or j['impl'] != 'reference'
    assert j['metric'] == 'pretrain_octuple_tokens_per_s' and j['unit'] == 'tokens/s' and j['higher_is_better'] is True
    assert j['n_gpus'] == 1 and j['steps'] == 20 and j['warmup'] == 3 and j['scaling'] == 'weak' and j['dtype'] == 'bf16'
    assert j['value'] > 1e5 and abs(j['value'] * j['ms_per_step'] / 1e3 - j['config']['global_batch'] * j['config']['seq_len']) < 1.0
    assert j['config']['workload'] and j['data'] == 'synthetic' and j['vs_baseline'] is None
    e = j['e2e']
    assert 0 < e['value'] <= 1.05 * j['value'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    r = j['roofline']
    assert r['bound'] in ('tensor', 'hbm') and r['unit'] in ('TFLOP/s', 'GB/s') and r['peak'] > 0
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-6 and 0.2 < r['frac'] < 1.0
    assert j['gpu_launches'] > 300 * 20
    c = j['clocks']
    assert isinstance(c['reasons'], list) and 'sm_mhz' in c and 'sm_max_mhz' in c
    assert c['sm_mhz'] is None or c['sm_max_mhz'] >= c['sm_mhz'] > 0      # (None: nvidia-smi gave no sample in the window)

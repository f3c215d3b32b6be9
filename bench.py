#!/usr/bin/env python
"""Benchmark of the PianoBART pretraining hot path on B200 (contract: see DESIGN.md section Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" is one full pretraining iteration of reference pretrain.py:120-209 on one batch of synthetic
Octuple tokens: noising -> forward (8+8 layer PianoBART, d=1024, S=1024) -> 8-head masked CE ->
backward -> gradient all-reduce (N>1) -> clip 3.0 -> AdamW.  Metric: Octuple tokens/s (whole job).

  value : inputs (original ids + noise plan) already resident in HBM when the timed region starts
  e2e   : through the public trainer call path with HOST buffers: host noise-plan generation, pinned H2D
          of ids + plan, the step, D2H of the loss/accuracy scalars - all inside the timed region
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_TOKEN = 1.29137e9  # fwd+bwd, causal attention counted as half


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if sm:
            sm_sorted = sorted(sm)
            out['sm_mhz'] = sm_sorted[len(sm_sorted) // 2]
            out['sm_max_mhz'] = max(mx)
            out['samples'] = len(sm)
        out['reasons'] = sorted(reasons)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def default_cfg():
    # reference defaults: pretrain.py:33-37, main.py:39-47
    return dict(d_model=1024, layers=8, heads=8, ffn=2048, max_pos=1024, seq=1024)


def cpu_baseline(batch=1, steps=2, warmup=1):
    """Times the CPU restatement of the path (oracle, 'port') on the host cores: fwd + loss + bwd."""
    import numpy as np
    import torch
    from oracle import params as P
    from oracle import pianobart_oracle as O
    c = default_cfg()
    torch.set_num_threads(os.cpu_count())
    cfg = O.Cfg(c['d_model'], c['layers'], c['layers'], c['heads'], c['ffn'], c['max_pos'])
    prm = P.make_params(c['d_model'], c['layers'], c['layers'], c['ffn'], c['max_pos'], 3)
    p = {k: torch.from_numpy(v).requires_grad_(True) for k, v in prm.items() if not k.startswith('decoder_linear')}
    p['decoder_linear.weight'], p['decoder_linear.bias'] = p['encoder_linear.weight'], p['encoder_linear.bias']
    S = c['seq']
    ori = torch.from_numpy(P.synth_ids(batch, S, 1234))
    dec = O.shift_right(ori, P.SOS)
    rs = np.random.RandomState(0)
    lm = torch.from_numpy((rs.rand(batch, S, 1) < 0.15).astype(np.float32).repeat(8, axis=2))
    keep = torch.ones(batch, S)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        h, _ = O.pianobart_forward(p, cfg, ori, dec, keep, keep)
        total, _ = O.pretrain_loss(O.lm_heads(p, h), ori, lm)
        for v in p.values():
            v.grad = None
        total.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per = sorted(times)[len(times) // 2]
    return {'value': batch * S / per, 'unit': 'tokens/s', 'cores': os.cpu_count(), 'kind': 'port',
            'sample': 'oracle fwd+loss+bwd fp32, default model, batch %d x seq %d, %d step(s) median, torch CPU %d threads'
                      % (batch, S, steps, os.cpu_count()), 's_per_step': per}


def decode_bench(lm, pb, dev, peaks, steps=192, warmup=16, batches=(1, 64)):
    """KV-cache decode (BASELINE.json configs[2]): encoder prompt 1024 tokens, batch 1 and 64; timed region = `steps`
    CUDA-graph replays (one generated Octuple token per sequence per replay), CUDA events on the launching stream."""
    import numpy as np
    import torch
    from oracle import params as P
    from pianobart_b200.generate import Generator
    out = {}
    d, L, F, S = 1024, 8, 2048, 1024
    w_bytes = 2 * (L * (4 * d * d + 2 * d * d + 2 * d * F) + 2048 * d + d * 1280)
    for B in batches:
        gen = Generator(lm, B, S, S)
        ids = torch.from_numpy(P.synth_ids(B, S, 4321)).to(dev)
        keep = torch.ones(B, S, device=dev)
        uni = np.random.RandomState(0).random_sample((B, S, 8))
        # teacher-force valid tokens so that no sequence stops early and the full decode work is timed
        forced = torch.from_numpy(P.synth_ids(B, S, 99)).to(dev)
        gen.start(ids, keep, uni, forced)
        gen.run_steps(warmup)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launches = gen.run_steps(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        byts = sum(w_bytes + B * (L * 2 * (t + 1) * d * 2 + L * 2 * S * d * 2) for t in range(warmup, warmup + steps))
        ach = byts / (ms / 1e3) / 1e9
        out['batch%d' % B] = {'tokens_per_s': B * steps / (ms / 1e3), 'us_per_step': ms / steps * 1e3,
                              'steps': steps, 'launches_per_step': gen.launches_per_step,
                              'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                           'frac': ach / peaks['hbm_gbs']}}
        del gen
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cb = cpu_baseline(batch=1, steps=max(1, min(args.steps, 3)), warmup=1 if args.warmup > 0 else 0)
    c = default_cfg()
    line = {'impl': 'reference', 'metric': 'pretrain_octuple_tokens_per_s', 'value': cb['value'], 'unit': 'tokens/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': cb['s_per_step'] * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'PianoBART pretrain step fwd+loss+bwd, default model d=1024 8+8 layers S=1024 '
                                   '(CPU, bounded sample batch 1)', 'seq_len': c['seq']},
            'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': cb['value'], 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=16, help='per-GPU batch (reference default pretrain.py:30)')
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-decode', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import numpy as np
    import torch
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        pg = dist.group.WORLD

    from oracle import params as P  # synthetic id generator only (numpy); not part of the timed path
    from pianobart_b200 import _lib as L
    from pianobart_b200.modules import BartConfig, PianoBart, PianoBartLM
    from pianobart_b200.pretrain import FusedAdamW, PretrainStep
    from pianobart_b200.vocab import build_octuple_vocab
    import random

    c = default_cfg()
    torch.manual_seed(2023)
    e2w, w2e = build_octuple_vocab()
    bc = BartConfig(max_position_embeddings=c['max_pos'], d_model=c['d_model'], encoder_layers=c['layers'],
                    decoder_layers=c['layers'], encoder_ffn_dim=c['ffn'], decoder_ffn_dim=c['ffn'],
                    encoder_attention_heads=c['heads'], decoder_attention_heads=c['heads'])
    dev = torch.device('cuda', local_rank)
    pb = PianoBart(bc, e2w, w2e, dtype=args.dtype)
    lm = PianoBartLM(pb).to(dev)
    lm.train()   # training-mode arithmetic: dropout(0.1) active as in the reference's Pretrainer.train()
    opt = FusedAdamW(pb, lr=2e-5, weight_decay=0.01)
    B, S = args.batch, c['seq']
    step = PretrainStep(lm, B, S, opt, 0.15, pg)
    lib = L.lib()
    random.seed(2023 + rank)
    np.random.seed(2023 + rank)
    nbatches = 4
    batches = [P.synth_ids(B, S, 1234 + 97 * rank + i) for i in range(nbatches)]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ("value")
    step.upload(batches[0])
    for _ in range(args.warmup):
        step.noise(); step.run(train=True)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    lib.pb_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step.noise(); step.run(train=True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(lib.pb_launch_count())
    clocks = sampler.stop() if sampler else None
    total, losses, accs = step.fetch_stats()

    # ---------------- end-to-end timing through host buffers ("e2e")
    # Input pipeline as a trainer runs it: while the GPU executes step i the host draws the noise plan of step i+1
    # and enqueues its pinned H2D copies behind step i (stream order keeps the device buffers consistent); the
    # loss/accuracy scalars of every step are read back (D2H + sync) before the next step is launched.
    for i in range(min(args.warmup, 2)):
        step.upload(batches[i % nbatches]); step.noise(); step.run(train=True); step.fetch_stats()
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    step.upload(batches[0])
    for i in range(args.steps):
        step.noise(); step.run(train=True)
        if i + 1 < args.steps:
            step.upload(batches[(i + 1) % nbatches])
        step.fetch_stats()
    e3.record()
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3 * 0.0)
    h2d, d2h = step.h2d_bytes, step.d2h_bytes

    # ---------------- dominant-kernel timing: every tcgen05 GEMM launch of one step bracketed by CUDA events
    gemm_flops, gemm_ms, gemm_n = 0.0, 0.0, 0
    if args.dtype == 'bf16':
        from pianobart_b200.engine import profile_gemms
        step.noise()
        gemm_flops, gemm_ms, gemm_n = profile_gemms(step)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t[0].item(), t[1].item()
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    tokens = world * B * S * args.steps
    value = tokens / (ms / 1e3)
    e2e_value = tokens / (ms_e2e / 1e3)
    per_gpu_tflops = value / world * FLOP_PER_TOKEN / 1e12
    line = {
        'metric': 'pretrain_octuple_tokens_per_s', 'value': value, 'unit': 'tokens/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': 'PianoBART span-mask/infill pretraining step (noise+fwd+loss+bwd+allreduce+clip+AdamW), '
                               'default model d=1024 8+8 layers 8 heads ffn 2048, synthetic Octuple ids',
                   'per_gpu_batch': B, 'global_batch': B * world, 'seq_len': S, 'parallelism': 'dp%d' % world,
                   'l2_policy': 'per-step working set (weights 0.35 GB + activations > 10 GB) is far larger than the 126 MB L2',
                   'dropout': 'p=%.2f applied (train mode, counter-based masks)' % pb.dropout_p()},
        'e2e': {'value': e2e_value, 'unit': 'tokens/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches,
        'clocks': clocks,
        'loss': total,
        'step_tflops_per_gpu': per_gpu_tflops,
        'step_frac_of_bf16_sustained': per_gpu_tflops / peaks['bf16_tflops_sustained'],
        'step_frac_of_bf16_burst': per_gpu_tflops / peaks['bf16_tflops'],
    }
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'r1_gemm_traffic.json')
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get('dram_bytes_per_launch_avg')
    if gemm_n:
        ach = gemm_flops / (gemm_ms / 1e3) / 1e12
        line['roofline'] = {'bound': 'tensor', 'kernel': 'gemm_tc_kernel (all %d tcgen05 GEMM launches of one step)' % gemm_n,
                            'achieved': ach, 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                            'frac': ach / peaks['bf16_tflops_sustained'], 'traffic': traffic,
                            'traffic_note': 'avg dram read+write bytes per GEMM launch from one ncu --set full capture '
                                            '(profiles/r1_gemm_traffic.json); algorithmic flops per launch avg = %.3e' % (gemm_flops / gemm_n),
                            'peak_source': peak_src + ' (sustained: kernel timed inside a long step)',
                            'share_of_step': gemm_ms / (ms / args.steps)}
    if world == 1 and not args.no_decode:
        try:
            line['decode'] = decode_bench(lm, pb, dev, peaks)
        except Exception as e:  # the headline metric must still be printed
            line['decode'] = {'error': repr(e)[:200]}
    if not args.no_cpu_baseline:
        cb = cpu_baseline(batch=1, steps=2, warmup=1)
        line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

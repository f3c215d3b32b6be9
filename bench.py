#!/usr/bin/env python
"""Benchmark of the PianoBART hot path on B200 (contract: see DESIGN.md section Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
                    [--workload pretrain|genft|seqcls|tokcls]

Default workload (BASELINE.json configs[1], the one the metric is quoted on): a "step" is one full pretraining iteration
of reference pretrain.py:120-209 on one batch of synthetic Octuple tokens: noising -> forward (8+8 layer PianoBART,
d=1024, S=1024) -> 8-head masked CE -> backward -> gradient all-reduce (N>1) -> clip 3.0 -> AdamW.
Metric: Octuple tokens/s (whole job).

  value : inputs (original ids + noise plan) already resident in HBM when the timed region starts
  e2e   : through the public trainer call (`Pretrainer.iteration`) with HOST batches: noise-plan generation (prefetch
          thread), pinned H2D of ids + plan, the step, D2H of the loss / accuracy scalars - all inside the timed region
The other workloads (configs[3], [4] and the generation finetune) time `FinetuneTrainer.step` / `GenerationTrainer.step`
the same way; KV-cache decode (configs[2]) is reported in the `decode` object of the default line.
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
FLOP_PER_TOKEN_HEADS = 3 * 2 * 1024 * 1280   # share of the 8 LM heads (absent from the classification workloads)


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


# ------------------------------------------------------------------------------------------------ CPU arm
def _find_reference():
    """The unmodified reference is pure Python: importable wherever its tree is present (/root/reference in the build
    container, baseline/_ref if a driver placed it there); it does not travel to the GPU box."""
    for p in (os.environ.get('PIANOBART_REF'), '/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if p and os.path.exists(os.path.join(p, 'PianoBart.py')) and os.path.exists(os.path.join(p, 'model.py')):
            return p
    return None


def cpu_baseline(batch=2, steps=2, warmup=1):
    """fwd + loss + bwd of the default model in train mode, fp32, batch 2 x 1024 synthetic Octuple
    ids, on all host cores.  kind 'reference': the UNMODIFIED reference (PianoBartLM + the loss of pretrain.py:112-118,
    179-189; harness shim transformers.AdamW = torch.optim.AdamW) when its tree is present; else kind 'port': the oracle
    restatement of the same arithmetic (the one place outside tests/ that may execute oracle/)."""
    import numpy as np
    import torch
    from oracle import params as P
    c = default_cfg()
    torch.set_num_threads(os.cpu_count())
    S = c['seq']
    ori = torch.from_numpy(P.synth_ids(batch, S, 1234))
    rs = np.random.RandomState(0)
    lm = torch.from_numpy((rs.rand(batch, S, 1) < 0.15).astype(np.float32).repeat(8, axis=2))
    keep = torch.ones(batch, S)
    ref = _find_reference()
    times = []
    if ref is not None:
        kind = 'reference'
        cwd = os.getcwd()
        sys.path.insert(0, ref)
        try:
            os.chdir(ref)
            import pickle
            import transformers
            transformers.AdamW = torch.optim.AdamW
            from transformers import BartConfig
            import PianoBart as ref_pb
            import model as ref_model
            with open(os.path.join(ref, 'Data', 'Octuple.pkl'), 'rb') as f:
                e2w, w2e = pickle.load(f)
            torch.manual_seed(2023)
            bc = BartConfig(max_position_embeddings=c['max_pos'], d_model=c['d_model'], encoder_layers=c['layers'],
                            decoder_layers=c['layers'], encoder_ffn_dim=c['ffn'], decoder_ffn_dim=c['ffn'],
                            encoder_attention_heads=c['heads'], decoder_attention_heads=c['heads'])
            pbm = ref_pb.PianoBart(bc, e2w, w2e)
            model = ref_model.PianoBartLM(pbm)
            model.train()
            dec = torch.empty_like(ori)
            dec[:, 1:] = ori[:, :-1]
            dec[:, 0] = torch.tensor(pbm.sos_word_np)
            loss_func = torch.nn.CrossEntropyLoss(reduction='none')
            n_tok = [len(pbm.e2w[k]) for k in pbm.e2w]
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                y = model.forward(ori, dec, keep, keep)
                losses = []
                for j in range(8):
                    l = loss_func(y[j].permute(0, 2, 1), ori[..., j]) * lm[:, :, j]
                    losses.append(torch.sum(l) / torch.sum(lm[:, :, j]))
                total = sum(a * b for a, b in zip(losses, n_tok)) / sum(n_tok)
                model.zero_grad()
                total.backward()
                dt = time.perf_counter() - t0
                if i >= warmup:
                    times.append(dt)
        finally:
            os.chdir(cwd)
            sys.path.remove(ref)
        what = 'unmodified reference PianoBartLM (train mode) fwd+loss+bwd'
    else:
        kind = 'port'
        from oracle import pianobart_oracle as O
        cfg = O.Cfg(c['d_model'], c['layers'], c['layers'], c['heads'], c['ffn'], c['max_pos'])
        prm = P.make_params(c['d_model'], c['layers'], c['layers'], c['ffn'], c['max_pos'], 3)
        p = {k: torch.from_numpy(v).requires_grad_(True) for k, v in prm.items() if not k.startswith('decoder_linear')}
        p['decoder_linear.weight'], p['decoder_linear.bias'] = p['encoder_linear.weight'], p['encoder_linear.bias']
        dec = O.shift_right(ori, P.SOS)
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
        what = 'oracle restatement (reference tree absent on this machine) fwd+loss+bwd'
    per = sorted(times)[len(times) // 2]
    return {'value': batch * S / per, 'unit': 'tokens/s', 'cores': os.cpu_count(), 'kind': kind,
            'sample': '%s, fp32, default model, batch %d x seq %d, %d timed step(s) (median) after %d warm-up, torch CPU %d threads'
                      % (what, batch, S, steps, warmup, os.cpu_count()), 's_per_step': per, 'steps_run': steps}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))          # bounded sample: a CPU step takes seconds
    warm = 1 if args.warmup > 0 else 0
    cb = cpu_baseline(batch=2, steps=steps, warmup=warm)
    c = default_cfg()
    line = {'impl': 'reference', 'metric': 'pretrain_octuple_tokens_per_s', 'value': cb['value'], 'unit': 'tokens/s',
            'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'steps_requested': args.steps,
            'ms_per_step': cb['s_per_step'] * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'PianoBART pretrain step fwd+loss+bwd, default model d=1024 8+8 layers S=1024 '
                                   '(CPU, bounded sample: batch 2, BASELINE.json configs[0])', 'seq_len': c['seq']},
            'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': cb['value'], 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ decode (configs[2])
def decode_bytes(B, t, S_enc, L=8, d=1024, F=2048):
    """decoder + front + head weights once per step, self K/V rows 0..t and cross K/V per sequence."""
    w = 2 * (L * (4 * d * d + 2 * d * d + 2 * d * F) + 2048 * d + d * 1280)
    return w + B * (L * 2 * (t + 1) * d * 2 + L * 2 * S_enc * d * 2)


def decode_bench(lm, pb, dev, peaks, steps=192, warmup=16, batches=(1, 64)):
    """KV-cache decode (BASELINE.json configs[2]): encoder prompt 1024 tokens, batch 1 and 64; timed region = `steps`
    generated Octuple tokens per sequence (batch 1: one persistent cooperative launch; batch 64: CUDA-graph replays),
    CUDA events on the launching stream."""
    import numpy as np
    import torch
    from oracle import params as P
    from pianobart_b200.generate import Generator
    out = {}
    S = 1024
    steps = int(os.environ.get('PIANOBART_B200_DECODE_STEPS', steps))
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
        byts = sum(decode_bytes(B, t, S) for t in range(warmup, warmup + steps))
        ach = byts / (ms / 1e3) / 1e9
        out['batch%d' % B] = {'tokens_per_s': B * steps / (ms / 1e3), 'us_per_step': ms / steps * 1e3,
                              'steps': steps, 'launches': launches,
                              'kernel': ('decode_persist_kernel (one cooperative launch)' if getattr(gen, 'persist', False) else
                                         'decode_batch_kernel (one cooperative launch)' if getattr(gen, 'persist_batch', False)
                                         else 'CUDA graph of per-op kernels'),
                              'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                           'frac': ach / peaks['hbm_gbs'], 'algorithmic_bytes_per_step': byts / steps}}
        del gen
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the reference default of the workload)')
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--workload', default='pretrain', choices=['pretrain', 'genft', 'seqcls', 'tokcls'])
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
    from pianobart_b200.modules import BartConfig, PianoBart
    from pianobart_b200.vocab import build_octuple_vocab
    import random

    c = default_cfg()
    e2w, w2e = build_octuple_vocab()
    bc = BartConfig(max_position_embeddings=c['max_pos'], d_model=c['d_model'], encoder_layers=c['layers'],
                    decoder_layers=c['layers'], encoder_ffn_dim=c['ffn'], decoder_ffn_dim=c['ffn'],
                    encoder_attention_heads=c['heads'], decoder_attention_heads=c['heads'])
    dev = torch.device('cuda', local_rank)
    torch.manual_seed(2023)
    pb = PianoBart(bc, e2w, w2e, dtype=args.dtype)          # same random-init weights on every rank
    torch.manual_seed(2023 ^ ((rank + 1) << 20))            # ... but rank-specific dropout masks (seed of the mask hash)
    S = c['seq']
    lib = L.lib()
    random.seed(2023 + rank)
    np.random.seed(2023 + rank)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload != 'pretrain':
        return finetune_workload(args, pb, dev, pg, rank, world, lib, barrier)

    from pianobart_b200.pretrain import Pretrainer
    B = args.batch or 16                         # reference default pretrain.py:30
    trainer = Pretrainer(pb, None, None, 2e-5, B, S, 0.15, False, [local_rank], process_group=pg, verbose=False)
    lm = trainer.model
    lm.train()   # training-mode arithmetic: dropout(0.1) active as in the reference's Pretrainer.train()
    step = trainer._step(B, S)
    nbatches = 4
    batches = [P.synth_ids(B, S, 1234 + 97 * rank + i) for i in range(nbatches)]

    # ---------------- device-resident timing ("value")
    step.upload(batches[0])
    for _ in range(args.warmup):
        step.noise(); step.run(train=True)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    lib.pb_reset_launch_count()
    from pianobart_b200 import pretrain as _PT
    _PT.GRAPH_REPLAYED_LAUNCHES[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step.noise(); step.run(train=True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # kernels of this library executed in the timed region: direct API launches + the kernel nodes of the captured step graph
    # that PretrainStep replays (the library's own counter only sees calls made outside a replay)
    api_launches = int(lib.pb_launch_count())
    launches = api_launches + int(_PT.GRAPH_REPLAYED_LAUNCHES[0])
    clocks = sampler.stop() if sampler else None
    total, losses, accs = step.fetch_stats()

    # ---------------- end-to-end timing through the public trainer API with host batches ("e2e")
    # Pretrainer.iteration: a host thread draws the noise plan of batch i+1 while the GPU runs step i; every step copies its
    # ids + plan from pinned memory and reads its 24 loss / accuracy scalars back (D2H + sync) before the next one starts.
    host_batches = [torch.from_numpy(batches[i % nbatches]) for i in range(args.steps)]
    trainer.iteration(host_batches[:min(max(args.warmup, 1), 2)], S, train=True)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    trainer.iteration(host_batches, S, train=True)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
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
                'ms_per_step': ms_e2e / args.steps, 'api': 'Pretrainer.iteration (host batches, plan prefetch thread)'},
        'gpu_launches': launches, 'gpu_launches_via_api_calls': api_launches,
        'clocks': clocks,
        'loss': total,
        'step_tflops_per_gpu': per_gpu_tflops,
        'step_frac_of_bf16_sustained': per_gpu_tflops / peaks['bf16_tflops_sustained'],
        'step_frac_of_bf16_burst': per_gpu_tflops / peaks['bf16_tflops'],
    }
    traffic, tsrc = None, None
    for name in ('r2_gemm_traffic.json', 'r1_gemm_traffic.json'):
        tp = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get('dram_bytes_per_launch_avg')
            tsrc = name
            break
    if gemm_n:
        ach = gemm_flops / (gemm_ms / 1e3) / 1e12
        line['roofline'] = {'bound': 'tensor', 'kernel': 'gemm_tc_kernel (all %d tcgen05 GEMM launches of one step)' % gemm_n,
                            'achieved': ach, 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                            'frac': ach / peaks['bf16_tflops_sustained'], 'frac_of_burst': ach / peaks['bf16_tflops'],
                            'traffic': traffic,
                            'traffic_note': 'avg dram read+write bytes per GEMM launch from one ncu --set full capture '
                                            '(profiles/%s); algorithmic flops per launch avg = %.3e' % (tsrc, gemm_flops / gemm_n),
                            'peak_source': peak_src + ' (sustained: kernel timed inside a multi-second step loop)',
                            'share_of_step': gemm_ms / (ms / args.steps)}
    if world == 1 and not args.no_decode:
        try:
            line['decode'] = decode_bench(lm, pb, dev, peaks)
        except Exception as e:  # the headline metric must still be printed
            line['decode'] = {'error': repr(e)[:200]}
    if world == 1 and not args.no_cpu_baseline:     # (rank 0 at N = 1 only: the other ranks of a multi-GPU run have left by now)
        cb = cpu_baseline(batch=2, steps=2, warmup=1)
        line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def finetune_workload(args, pb, dev, pg, rank, world, lib, barrier):
    """BASELINE.json configs[3] (sequence classification, B=8, 8 classes), configs[4] (token classification, velocity: 7+1
    classes through the replacement decoder front end) and the generation finetune (B=8): one optimisation step of the
    trainer's public `step` with HOST batches (H2D + loss D2H inside the timed region) = e2e; `value` = the same step with
    the batch already on the device."""
    import numpy as np
    import torch
    from oracle import params as P
    c = default_cfg()
    S = c['seq']
    B = args.batch or 8                          # finetune.py:33 / finetune_generation.py:29
    ids = torch.from_numpy(P.synth_ids(B, S, 777 + rank))
    rs = np.random.RandomState(5 + rank)
    if args.workload == 'genft':
        from pianobart_b200.finetune_generation import GenerationTrainer
        tr = GenerationTrainer(pb.to(dev), None, None, None, 2e-6, None, False, [dev.index], process_group=pg, verbose=False)
        y = torch.from_numpy(P.synth_ids(B, S, 888 + rank))
        tr.model.train()
        run_host = lambda: tr.step(ids, y, train=True)
        ids_d, y_d = ids.to(dev), y.to(dev)
        run_dev = lambda: tr.step(ids_d, y_d, train=True)
        name = 'generation finetune step (GenerationTrainer.step, y_shift = x)'
        flop = FLOP_PER_TOKEN
    else:
        from pianobart_b200.finetune import FinetuneTrainer
        seq = args.workload == 'seqcls'
        cn = 8 if seq else 7
        tr = FinetuneTrainer(pb, None, None, None, 2e-5, cn, c['d_model'], None, False, [dev.index], SeqClass=seq, process_group=pg)
        y = torch.from_numpy(rs.randint(0, cn, size=(B,) if seq else (B, S)))
        tr.model.train()
        run_host = lambda: tr.step(ids, y, mode=0)[0].item()
        ids_d, y_d = ids.to(dev), y.to(dev)
        run_dev = lambda: tr.step(ids_d, y_d, mode=0)
        name = ('composer sequence classification finetune step (8 classes)' if seq
                else 'velocity token classification finetune step (7+1 classes, label-embedding decoder front end)')
        flop = FLOP_PER_TOKEN - FLOP_PER_TOKEN_HEADS
    warm = max(args.warmup, 3)
    for _ in range(warm):
        run_dev()
    barrier()
    lib.pb_reset_launch_count()
    from pianobart_b200 import pretrain as _PT
    _PT.GRAPH_REPLAYED_LAUNCHES[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_dev()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(lib.pb_launch_count()) + int(_PT.GRAPH_REPLAYED_LAUNCHES[0])
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        run_host()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t[0].item(), t[1].item()
    if rank == 0:
        peaks, _ = load_peaks()
        tokens = world * B * S * args.steps
        value = tokens / (ms / 1e3)
        tf = value / world * flop / 1e12
        print(json.dumps({
            'metric': 'finetune_octuple_tokens_per_s', 'value': value, 'unit': 'tokens/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': warm, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': name + ', default model, train mode', 'per_gpu_batch': B, 'seq_len': S,
                       'parallelism': 'dp%d' % world},
            'e2e': {'value': tokens / (ms_e2e / 1e3), 'unit': 'tokens/s', 'h2d_bytes_per_step': int(ids.numel() * 8 + y.numel() * 8),
                    'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'step_tflops_per_gpu': tf, 'step_frac_of_bf16_sustained': tf / peaks['bf16_tflops_sustained']}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

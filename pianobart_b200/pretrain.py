"""Pretraining hot loop, mirroring reference pretrain.py (`Pretrainer`, pretrain.py:51-546).

`Pretrainer` keeps the reference's constructor, train()/valid()/iteration()/gen_mask()/
compute_loss()/save_checkpoint() interface and printed metrics, but one iteration is a single
fused device step (no per-sample D2H/H2D, no host argmax):

  host : noise plan for the batch (noising.py, bit-exact RNG)      pretrain.py:131-144
  H2D  : original ids (int16) + plan, from pinned staging buffers
  dev  : pb_noise_apply -> enc/dec ids, targets, loss mask, pad masks   pretrain.py:128-153
         pb_mask_sums  (-> all-reduce of the 8 denominators under data parallelism)
         forward plan (front end, 8+8 BART layers, heads as one N=1280 GEMM)
         pb_heads_ce   masked CE x8 + weights + argmax accuracy + dlogits  pretrain.py:163-189
         backward plan (gradient buckets all-reduced on a side stream as they become final)
         pb_sumsq + pb_adamw (clip 3.0 + HF AdamW, refreshes the bf16 working weights)  pretrain.py:192-196
"""
import ctypes as C
import os
import queue
import shutil
import sys
import threading

import numpy as np
import torch

from . import _lib as L
from . import engine as E
from . import noising
from .modules import PianoBart, PianoBartLM

# e2w pickle key order -> the n_tok sequence pretrain.py:184-189 multiplies the (classes-ordered) losses with
LOSS_WEIGHTS = [262, 134, 262, 134, 38, 135, 55, 260]


# kernels executed by replaying captured step graphs (the library's launch counter only sees direct API calls); bench.py adds
# this to pb_launch_count() for its gpu_launches figure
GRAPH_REPLAYED_LAUNCHES = [0]


class FusedAdamW:
    """HF-semantics AdamW (transformers 4.29 `AdamW(lr, weight_decay=0.01)`: betas (0.9, 0.999), eps 1e-6,
    bias correction folded in the step size, decay after the update) over the flat parameter buffer,
    with torch `clip_grad_norm_` folded in.  One launch for the norm, two for the update."""

    def __init__(self, pianobart, lr, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-6, max_grad_norm=3.0):
        self.pb = pianobart
        self.lr, self.wd, self.betas, self.eps, self.max_grad_norm = lr, weight_decay, betas, eps, max_grad_norm
        self.step_count = 0
        self.m = self.v = None
        self.gnorm_sq = None

    def _ensure(self):
        pb = self.pb
        pb._ensure_packed()
        if self.m is None or self.m.device != pb._flat.device or self.m.numel() != pb._flat.numel():
            self.m = torch.zeros_like(pb._flat)
            self.v = torch.zeros_like(pb._flat)
            self.gnorm_sq = torch.zeros(1, device=pb._flat.device, dtype=torch.float32)

    def step(self, grad_scale=1.0, gnorm=None):
        """gnorm: a device fp32 scalar that already holds the squared norm of the (all-reduced) gradient buffer - the
        partials the training step accumulates range by range during backward (engine.Plan.grads_final); None: one pass
        over the buffer here."""
        self._ensure()
        pb, lib, s = self.pb, L.lib(), L.stream_ptr()
        lay = pb.layout
        n = lay.size
        self.step_count += 1
        if gnorm is None:
            gnorm = self.gnorm_sq
            gnorm.zero_()
            L.check(lib.pb_sumsq(C.c_void_p(pb._grad.data_ptr()), C.c_longlong(n), C.c_void_p(gnorm.data_ptr()), s), 'sumsq')
        bf16 = pb.pb_dtype == E.PB_BF16
        for lo, hi, scale in ((0, lay.emb_end, 16.0), (lay.emb_end, n, 1.0)):
            wp = C.c_void_p(pb._wact.data_ptr() + lo * 2) if bf16 else C.c_void_p(None)
            L.check(lib.pb_adamw(C.c_void_p(pb._flat.data_ptr() + lo * 4), C.c_void_p(self.m.data_ptr() + lo * 4),
                                 C.c_void_p(self.v.data_ptr() + lo * 4), C.c_void_p(pb._grad.data_ptr() + lo * 4), wp,
                                 C.c_longlong(hi - lo), C.c_float(self.lr), C.c_float(self.betas[0]),
                                 C.c_float(self.betas[1]), C.c_float(self.eps), C.c_float(self.wd), self.step_count,
                                 C.c_void_p(gnorm.data_ptr()), C.c_float(self.max_grad_norm),
                                 C.c_float(grad_scale), C.c_float(scale), s), 'adamw')
        if not bf16:
            pb.mark_weights_dirty()
            pb._sync_weights()

    def state_dict(self):
        self._ensure()
        return {'step': self.step_count, 'exp_avg': self.m, 'exp_avg_sq': self.v, 'lr': self.lr,
                'weight_decay': self.wd, 'betas': self.betas, 'eps': self.eps}

    def load_state_dict(self, sd):
        self._ensure()
        self.step_count = sd['step']
        self.m.copy_(sd['exp_avg'])
        self.v.copy_(sd['exp_avg_sq'])


class PretrainStep:
    """One fused device step for a fixed (batch, seq) shape."""

    def __init__(self, lm, B, S, optimizer=None, mask_percent=0.15, process_group=None, dropout=None,
                 loss_weights=None, loss_norm=None):
        """dropout: None = follow lm.training (HF config.dropout in train() mode, 0 in eval()), or an explicit p.
        loss_weights / loss_norm: total = sum_i w_i L_i / loss_norm (default: pretrain.py:184-189, w = n_tok in e2w key
        order, norm = sum(n_tok); finetune_generation.py:238-250 passes w_i = extra_i * n_tok_i with the same norm)."""
        self.lm, self.pb = lm, lm.pianobart
        pb = self.pb
        pb._ensure_packed()
        self.B, self.S = B, S
        self.mask_percent = mask_percent
        self.opt = optimizer
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        dev = pb._flat.device
        self.dev = dev
        self.lib = L.lib()
        self.graph = pb._graph(B, S, S, True, True, dropout)
        self._pack_gen = pb._pack_gen
        self._cg, self._cg_n, self._graph_ok, self._eager_runs = None, 0, True, 0
        # clip norm from partials accumulated behind the 'grads_final' markers of the backward plan / the all-reduce of each
        # gradient bucket instead of a separate pass over the 0.7 GB gradient buffer after backward - only when
        # those ranges tile the whole flat buffer exactly once.  Measured (profiles/r2_summary.md section 16): data parallel
        # 29.85 vs 29.91 ms per step (on), single GPU 28.69 vs 28.62 ms (the partials on the side stream compete with the
        # power-capped GEMMs for as long as the separate pass takes) - so the default is on for world > 1 only;
        # PIANOBART_B200_NORM_PARTIALS=1|0 forces it.
        self._norm_partials = False
        # north-star fusion 3 (bf16 mode): MLM heads + masked CE in one kernel (csrc/heads_ce_tc.cu)
        self.fused_ce = (pb.pb_dtype == E.PB_BF16 and self.graph.d % 64 == 0
                         and os.environ.get('PIANOBART_B200_FUSED_CE', '1') != '0')
        M = B * S
        self.M = M
        # pinned host staging (two sets: the trainer stages batch i+1 while the copies of batch i may still be queued)
        # + device inputs
        self._stage = [dict(ori=torch.empty(B, S, 8, dtype=torch.int16).pin_memory(),
                            src=torch.empty(B, S, dtype=torch.int32).pin_memory(),
                            loss=torch.empty(B, S, dtype=torch.uint8).pin_memory(),
                            mode=torch.empty(B, dtype=torch.int32).pin_memory(),
                            rand=torch.zeros(max(1, M), 8, dtype=torch.int32).pin_memory()) for _ in range(2)]
        self._stage_i = 0
        self.d_ori = torch.empty(B, S, 8, dtype=torch.int16, device=dev)
        self.d_src = torch.empty(B, S, dtype=torch.int32, device=dev)
        self.d_loss = torch.empty(B, S, dtype=torch.uint8, device=dev)
        self.d_mode = torch.empty(B, dtype=torch.int32, device=dev)
        self.d_rand = torch.zeros(max(1, M), 8, dtype=torch.int32, device=dev)
        self.targets = torch.empty(M * 8, dtype=torch.int32, device=dev)
        self.loss_mask = torch.empty(M * 8, dtype=torch.float32, device=dev)
        # stats: [0:8] loss numerators, [8:16] correct counts, [16:24] mask sums (denominators)
        self.stats = torch.zeros(24, dtype=torch.float32, device=dev)
        self.h_stats2 = [torch.zeros(24, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._stats_ev = [torch.cuda.Event() for _ in range(2)]
        self._stats_slot = 0
        self.pad = (C.c_int * 8)(*[int(x) for x in pb.pad_word_np])
        self.mask = (C.c_int * 8)(*[int(x) for x in pb.mask_word_np])
        self.sos = (C.c_int * 8)(*[int(x) for x in pb.sos_word_np])
        self.seg = (C.c_int * 8)(*E.N_TOKENS)
        self.loss_weights = [float(x) for x in (loss_weights or LOSS_WEIGHTS)]
        self.loss_norm = float(loss_norm if loss_norm is not None else sum(self.loss_weights))
        self.w = (C.c_float * 8)(*self.loss_weights)
        # the kernel normalises by sum(w); rescale when the caller's normaliser differs
        self.grad_scale = sum(self.loss_weights) / self.loss_norm
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        if self.graph.bwd is not None and os.environ.get('PIANOBART_B200_NORM_PARTIALS', '1' if self.world > 1 else '0') != '0':
            rng = sorted((a[1], a[2]) for name, fn, a in self.graph.bwd.ops if name == 'marker' and a and a[0] == 'grads_final')
            pos = 0
            for lo, hi in rng:
                if lo != pos:
                    break
                pos = hi
            else:
                self._norm_partials = (pos == pb.layout.size and hasattr(self.graph, 'gnorm'))
        self.launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 24 * 4

    # -- stage 1: host plan + H2D + noising kernel
    def upload(self, ori_batch, choices=None, plan=None):
        """ori_batch: (B,S,8) integer tensor or array on the HOST.  plan: a NoisePlan drawn ahead of time for exactly
        this batch (PlanPrefetcher); otherwise it is drawn here."""
        ori = ori_batch.numpy() if isinstance(ori_batch, torch.Tensor) else np.asarray(ori_batch)
        if plan is None:
            plan = noising.make_plan(ori, self.S, self.mask_percent, choices)
        h = self._stage[self._stage_i]
        self._stage_i ^= 1
        h['ori'].numpy()[...] = ori
        h['src'].numpy()[...] = plan.src
        h['loss'].numpy()[...] = plan.loss
        h['mode'].numpy()[...] = plan.loss_mode
        nr = len(plan.rand_tok)
        if nr:
            h['rand'].numpy()[:nr] = np.asarray(plan.rand_tok, dtype=np.int32)
        # stream-ordered: these copies run after every kernel already queued, i.e. after the previous step has consumed
        # the device input buffers
        self.d_ori.copy_(h['ori'], non_blocking=True)
        self.d_src.copy_(h['src'], non_blocking=True)
        self.d_loss.copy_(h['loss'], non_blocking=True)
        self.d_mode.copy_(h['mode'], non_blocking=True)
        if nr:
            self.d_rand[:nr].copy_(h['rand'][:nr], non_blocking=True)
        self.h2d_bytes = (h['ori'].numel() * 2 + h['src'].numel() * 4 + h['loss'].numel() + h['mode'].numel() * 4 + nr * 32)
        return plan

    def noise(self):
        g, s = self.graph, L.stream_ptr()
        P = C.c_void_p
        L.check(self.lib.pb_noise_apply(P(self.d_ori.data_ptr()), P(self.d_src.data_ptr()), P(self.d_rand.data_ptr()),
                                        P(self.d_loss.data_ptr()), P(self.d_mode.data_ptr()), P(g.enc_ids.data_ptr()),
                                        P(g.dec_ids.data_ptr()), P(self.targets.data_ptr()), P(self.loss_mask.data_ptr()),
                                        P(g.enc_keep.data_ptr()), P(g.dec_keep.data_ptr()), self.B, self.S, self.pad,
                                        self.mask, self.sos, s), 'noise_apply')
        self.launches += 1

    def set_device_batch(self, enc_ids, dec_ids, targets, loss_mask, enc_keep, dec_keep):
        """Alternative to upload()+noise(): already-noised device tensors (parity tests, generation finetune)."""
        g = self.graph
        g.set_inputs(enc_ids, enc_keep, dec_ids, dec_keep)
        self.targets.copy_(targets.reshape(-1))
        self.loss_mask.copy_(loss_mask.reshape(-1))

    # -- stage 2: forward + loss (+ backward + optimizer)
    def run(self, train=True, profile=None):
        pb = self.pb
        pb.check_pack_generation(self._pack_gen, 'PretrainStep')
        pb._sync_weights()
        pb._live_graph = None  # the fused step owns the graph buffers; autograd must not reuse them
        # The ~330 launches of forward + loss + backward (incl. the side-stream fork / joins) replay as ONE CUDA graph once the
        # step is warm (single GPU, training mode): same kernels, same order, programmatic-dependent-launch edges kept.
        # The optimizer stays outside (its bias-correction scalars change every step).  PIANOBART_B200_STEP_GRAPH=0: eager.
        # (Data parallel steps stay eager: capturing the NCCL all-reduces works and is 0.3 ms faster at 2 GPUs, but the
        # processes then hang in teardown until killed - profiles/r2_summary.md section 12.)
        if (train and profile is None and self.world == 1 and self._graph_ok
                and os.environ.get('PIANOBART_B200_STEP_GRAPH', '1') != '0'):
            if self._cg is None and self._eager_runs >= 2:
                try:
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg):
                        self._cg_n = self._enqueue(True, None)
                    self._cg = cg
                except Exception as e:            # capture unsupported for some node: keep launching eagerly (same kernels)
                    self._graph_ok = False
                    torch.cuda.synchronize()
                    sys.stderr.write('pianobart_b200: step graph capture failed (%s); launching eagerly\n' % (str(e)[:200],))
            if self._cg is not None:
                self._cg.replay()
                n = self._cg_n
                GRAPH_REPLAYED_LAUNCHES[0] += n
                if self.opt is not None:
                    n += self._opt_step()
                self.launches += n
                return n
        self._eager_runs += 1
        n = self._enqueue(train, profile)
        if train and self.opt is not None:
            n += self._opt_step()
        self.launches += n
        return n

    def _enqueue(self, train, profile):
        """Queues forward + loss (+ backward) of one step on the current stream (and the graph's side / comm streams)."""
        g, lib, s = self.graph, self.lib, L.stream_ptr()
        pb = self.pb
        P = C.c_void_p
        self.stats.zero_()
        M = self.M
        L.check(lib.pb_mask_sums(P(self.loss_mask.data_ptr()), P(self.stats.data_ptr() + 64), C.c_longlong(M), 8, s), 'mask_sums')
        den_ready = None
        if self.world > 1:
            # global denominators (pretrain.py:117 on the full batch): the all-reduce runs on the comm stream during the
            # forward pass - only the loss kernel at its end needs them (on the main stream its latency sat between steps)
            import torch.distributed as dist
            cur = torch.cuda.current_stream()
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.stats[16:24], group=self.pg)
                den_ready = torch.cuda.Event()
                den_ready.record()
        side = g.side_stream() if train else None
        if side is not None:
            # the 695 MB gradient buffer is cleared on the side stream while the forward pass runs (nothing reads or writes
            # it before backward; the previous step's optimizer, its last reader, precedes this point on the main stream)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                L.check(lib.pb_fill_zero(P(pb._grad.data_ptr()), C.c_longlong(pb._grad.numel() * 4), L.stream_ptr()), 'fill_zero')
        n = g.fwd.run(profile=profile)
        if den_ready is not None:
            torch.cuda.current_stream().wait_event(den_ready)
        if self.fused_ce:
            # heads GEMM + masked CE + accuracy + dlogits in one tcgen05 kernel: the fp32 logits never reach HBM
            L.check(lib.pb_heads_ce_fused(P(g.out.data_ptr()), C.c_longlong(g.d), P(g.W('heads.w')), P(g.Pf('heads.b')),
                                          P(self.targets.data_ptr()), P(self.loss_mask.data_ptr()),
                                          P(self.stats.data_ptr() + 64), P(self.stats.data_ptr()),
                                          P(self.stats.data_ptr() + 32), P(g.dlogits.data_ptr()) if train else P(None),
                                          P(None), C.c_longlong(M), g.d, 8, self.seg, self.w, C.c_float(self.grad_scale),
                                          L.stream_ptr()),
                    'heads_ce_fused')
            n += 1
        else:
            n += g.heads_plan().run(profile=profile)
            L.check(lib.pb_heads_ce(P(g.logits.data_ptr()), P(self.targets.data_ptr()), P(self.loss_mask.data_ptr()),
                                    P(self.stats.data_ptr() + 64), P(self.stats.data_ptr()), P(self.stats.data_ptr() + 32),
                                    P(g.dlogits.data_ptr()) if train else P(None), P(None), C.c_longlong(M), 8, self.seg,
                                    self.w, C.c_float(self.grad_scale), pb.pb_dtype, L.stream_ptr()), 'heads_ce')
            n += 2
        if train:
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)
            else:
                pb._grad.zero_()
            if self.world > 1:
                from .parallel import BucketReducer
                red = BucketReducer(pb._grad, self.pg, comm_stream=self.comm_stream,
                                    after_reduce=self._bucket_norm if self._norm_partials else None)
                n += g.bwd.run(on_marker=red.on_final, profile=profile, side_stream=g.side_stream(), on_op=red.on_attention,
                               norm_partials='zero_only' if self._norm_partials else None)
                red.finish()
            else:
                n += g.bwd.run(profile=profile, side_stream=g.side_stream(),
                               norm_partials='local' if self._norm_partials else None)
        return n

    def _bucket_norm(self, lo, hi):
        """Data parallel: squared norm of an all-reduced bucket, queued behind its all-reduce on the comm stream."""
        L.check(self.lib.pb_sumsq(C.c_void_p(self.pb._grad.data_ptr() + lo * 4), C.c_longlong(hi - lo),
                                  C.c_void_p(self.graph.gnorm.data_ptr()), L.stream_ptr()), 'sumsq')

    def _opt_step(self):
        self.opt.step(gnorm=self.graph.gnorm) if self._norm_partials else self.opt.step()
        return 2 if self._norm_partials else 3

    def queue_stats(self):
        """Queues the D2H copy of the step's 24 scalars (behind the step, on the current stream) into one of two pinned
        slots and returns a ticket for collect_stats(): the trainer reads step i back only after step i + 1 has been
        launched, so the device never idles between steps waiting for the host."""
        k = self._stats_slot
        self._stats_slot ^= 1
        if self.world > 1:
            # numerators / correct counts summed over the ranks on the comm stream: the next step does not wait for it
            import torch.distributed as dist
            st = self.stats.clone()
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(st[0:16], group=self.pg)
                self.h_stats2[k].copy_(st, non_blocking=True)
                self._stats_ev[k].record()
            st.record_stream(self.comm_stream)
            return k
        self.h_stats2[k].copy_(self.stats, non_blocking=True)
        self._stats_ev[k].record()
        return k

    def collect_stats(self, ticket):
        """-> (total_loss, losses[8], accs[8]) like pretrain.py:171-189, for the step queue_stats() was called after."""
        self._stats_ev[ticket].synchronize()
        v = self.h_stats2[ticket].numpy().astype(np.float64)
        num, cor, den = v[0:8], v[8:16], v[16:24]
        with np.errstate(divide='ignore', invalid='ignore'):
            losses = num / den
            accs = cor / den
        total = float(np.sum(losses * np.array(self.loss_weights)) / self.loss_norm)
        return total, losses, accs

    def fetch_stats(self):
        """D2H of the 24 step scalars + synchronisation; returns (total_loss, losses[8], accs[8])."""
        return self.collect_stats(self.queue_stats())


class PlanPrefetcher:
    """Input pipeline stage: a host thread pulls the next batch from the data iterator and draws its noise
    plan while the GPU executes the current step.  Only this thread touches the Python / numpy RNG streams during an
    iteration and it handles the batches strictly in order, so the corruption decisions are the ones a sequential run
    (and the reference, pretrain.py:131-144) draws.  Under data parallelism with `world > 1` the iterator yields the GLOBAL
    batch on every rank; the plan is drawn for all of it and sliced to this rank's samples (one logical RNG stream)."""

    def __init__(self, data_iter, mask_percent, rank=0, world=1, depth=2):
        self.it, self.mask_percent, self.rank, self.world = data_iter, mask_percent, rank, world
        self.q = queue.Queue(maxsize=depth)
        # The plan generator is pure Python: while it runs it holds the GIL, and the trainer thread returning from its
        # per-step device synchronisation would wait up to one interpreter switch interval (5 ms by default = 15 % of a
        # 31 ms step, measured) before it can launch the next step.  A short interval bounds that bubble to ~0.2 ms.
        self._old_switch = sys.getswitchinterval()
        sys.setswitchinterval(min(self._old_switch, 2e-4))
        self.t = threading.Thread(target=self._work, daemon=True)
        self.t.start()

    def _work(self):
        try:
            for batch in self.it:
                ori = batch.numpy() if isinstance(batch, torch.Tensor) else np.asarray(batch)
                plan = noising.make_plan(ori, ori.shape[1], self.mask_percent)
                if self.world > 1:
                    per = ori.shape[0] // self.world
                    lo = self.rank * per
                    plan = noising.slice_plan(plan, lo, lo + per)
                    ori = ori[lo:lo + per]
                self.q.put((np.ascontiguousarray(ori), plan))
            self.q.put(None)
        except BaseException as e:     # surfaced in the consumer thread
            self.q.put(e)
        finally:
            sys.setswitchinterval(self._old_switch)

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            yield item


class Pretrainer:
    """Same interface as reference `Pretrainer` (pretrain.py:51-209)."""

    def __init__(self, pianobart: PianoBart, train_dataloader, valid_dataloader, lr, batch, max_seq_len, mask_percent,
                 cpu, cuda_devices=None, process_group=None, verbose=True, global_batches=False):
        """global_batches (data parallel only): the data loaders yield the GLOBAL batch on every rank; each rank draws the
        noise plan for all of it in global sample order and trains on its slice, so a seeded run makes the corruption
        decisions of the single-process reference.  Otherwise every rank noises its own batches from
        its own RNG streams."""
        if cpu or not torch.cuda.is_available():
            raise L.PBError('pianobart_b200.Pretrainer has no CPU path (sm_100a kernels only)')
        dev = 'cuda'
        if process_group is None and cuda_devices is not None and len(cuda_devices) >= 1:
            dev += ':' + str(cuda_devices[0])
        elif process_group is not None:
            dev += ':' + str(torch.cuda.current_device())
        self.device = torch.device(dev)
        self.pianobart = pianobart.to(self.device)
        self.model = PianoBartLM(pianobart).to(self.device)
        self.total_params = sum(p.numel() for p in self.model.parameters() if p.requires_grad)
        self.verbose = verbose
        if verbose:
            print('# total parameters:', self.total_params)
        if cuda_devices is not None and len(cuda_devices) > 1 and process_group is None and verbose:
            print('pianobart_b200: multi-GPU data parallelism is one process per GPU (torchrun); '
                  'this process uses %s only' % dev)
        self.train_data, self.valid_data = train_dataloader, valid_dataloader
        self.optim = FusedAdamW(self.pianobart, lr=lr, weight_decay=0.01)
        self.batch, self.max_seq_len, self.mask_percent = batch, max_seq_len, mask_percent
        self.pg = process_group
        self.rank, self.world = 0, 1
        if process_group is not None:
            import torch.distributed as dist
            self.rank, self.world = dist.get_rank(process_group), dist.get_world_size(process_group)
        self.global_batches = bool(global_batches) and self.world > 1
        self.prefetch = os.environ.get('PIANOBART_B200_PREFETCH', '1') != '0'
        self._steps = {}
        self._train_mode = True

    def _step(self, B, S):
        k = (B, S, self.model.training)
        if k not in self._steps:
            self._steps[k] = PretrainStep(self.model, B, S, self.optim, self.mask_percent, self.pg)
        return self._steps[k]

    def train(self):
        self.model.train()
        return self.iteration(self.train_data, self.max_seq_len)

    def valid(self):
        self.model.eval()
        return self.iteration(self.valid_data, self.max_seq_len, train=False)

    def compute_loss(self, predict, target, loss_mask):
        """pretrain.py:112-118 on torch tensors (kept for API compatibility; the fused step does not call it)."""
        loss = torch.nn.functional.cross_entropy(predict, target, reduction='none') * loss_mask
        return torch.sum(loss) / torch.sum(loss_mask)

    def gen_mask(self, input_ids, choice=None):
        """pretrain.py:211-546 for one (S,8) sample: returns (noised ids, loss mask) as CPU tensors, consuming
        the RNG streams like the reference.  Data movement runs on the device kernel."""
        ori = input_ids.cpu().numpy()[None]
        st = self._step(1, ori.shape[1])
        st.upload(ori, None if choice is None else [choice])
        st.noise()
        enc = st.graph.enc_ids.view(1, -1, 8)[0].cpu().long()
        lm = st.loss_mask.view(-1, 8).cpu()
        return enc, lm

    def _batches(self, training_data):
        """(ori, plan) pairs in order: drawn one step ahead on a host thread (default), or inline."""
        world = self.world if self.global_batches else 1
        if self.prefetch:
            return iter(PlanPrefetcher(iter(training_data), self.mask_percent, self.rank, world))

        def inline():
            for batch in training_data:
                ori = batch.numpy() if isinstance(batch, torch.Tensor) else np.asarray(batch)
                plan = noising.make_plan(ori, ori.shape[1], self.mask_percent)
                if world > 1:
                    per = ori.shape[0] // world
                    plan = noising.slice_plan(plan, self.rank * per, (self.rank + 1) * per)
                    ori = ori[self.rank * per:(self.rank + 1) * per]
                yield ori, plan
        return inline()

    def iteration(self, training_data, max_seq_len, train=True):
        total_acc, total_losses, nb = np.zeros(8), 0.0, 0
        it = self._batches(training_data)

        def stage(item):
            """Host -> pinned -> device copies of one batch, queued behind whatever the stream is already running."""
            if item is None:
                return None
            ori, plan = item
            st = self._step(ori.shape[0], ori.shape[1])
            st.upload(ori, plan=plan)
            return st
        def collect(pending):
            nonlocal total_losses, total_acc, nb
            total, losses, accs = pending[0].collect_stats(pending[1])
            if self.verbose:
                sys.stdout.write('Loss: {:06f} | loss: {:03f}, {:03f}, {:03f}, {:03f}, {:03f}, {:03f}, {:03f}, {:03f}\n'.format(total, *losses))
                sys.stdout.write('Acc: {:06f} | acc: {:03f}, {:03f}, {:03f}, {:03f}, {:03f}, {:03f}, {:03f}, {:03f}\n'.format(np.average(accs), *accs))
            total_losses += total
            total_acc += accs
            nb += 1
        st = stage(next(it, None))
        pending = None
        while st is not None:
            st.noise()
            st.run(train=train)
            ticket = st.queue_stats()
            # Software pipeline, two steps deep: step i is on the device; the host now reads back the scalars of step i - 1
            # (its wait ends while step i still runs, so the device never idles between steps), then stages batch i + 1 (plan
            # drawn by the prefetch thread; the pinned set it reuses was last read by the copies of batch i - 1, complete by
            # then) and queues its H2D copies behind step i.
            if pending is not None:
                collect(pending)
            pending = (st, ticket)
            st = stage(next(it, None))
        if pending is not None:
            collect(pending)
        nb = max(nb, 1)
        return round(total_losses / nb, 3), [round(float(x) / nb, 3) for x in total_acc]

    def save_checkpoint(self, epoch, best_acc, valid_acc, valid_loss, train_loss, is_best, filename):
        """Same dict layout as pretrain.py:96-110 ('state_dict' = PianoBart only, reference key names)."""
        state = {'epoch': epoch + 1,
                 'state_dict': {k: v.detach().clone() for k, v in self.pianobart.state_dict().items()},
                 'best_acc': best_acc, 'valid_acc': valid_acc, 'valid_loss': valid_loss, 'train_loss': train_loss,
                 'optimizer': self.optim.state_dict()}
        torch.save(state, filename)
        best_mdl = filename.split('.')[0] + '_best.ckpt'
        if is_best:
            shutil.copyfile(filename, best_mdl)

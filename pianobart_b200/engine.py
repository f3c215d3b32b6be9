"""Execution engine: builds, for one (batch, seq) shape, the exact sequence of C-ABI kernel
launches that computes the PianoBART backbone forward and backward, and replays it.

Host side of the hot path: the arithmetic of
PianoBart.forward (reference PianoBart.py:56-80) + HF BartModel (modeling_bart.py) + MLM heads
(model.py:119-126) + masked CE (pretrain.py:112-118,179-189), expressed as a flat "plan" of
ctypes calls into libpianobart_b200.so.  PyTorch is used for device memory and streams only;
no torch operator computes anything on this path, and there is no fallback: the plan consists
solely of library launches.

Two dtype modes share the same plan:
    fp32  - SIMT fp32 GEMM (pb_gemm_f32); parity mode (loss <= 1e-4 rel. vs the reference)
    bf16  - tcgen05/TMEM/TMA GEMM (pb_gemm_bf16); production mode
"""
import ctypes as C
import math
import os

import torch

from . import _lib as L

PB_F32, PB_BF16 = 0, 1
N_TOKENS = [262, 134, 135, 262, 134, 38, 260, 55]
VOCAB = sum(N_TOKENS)  # 1280


def _ptr(t, off_elems=0):
    return t.data_ptr() + off_elems * t.element_size()


class ParamLayout:
    """Flat fp32 parameter layout.  Reference state_dict tensors are views
    into one flat buffer arranged so that fused operands are contiguous:
    q_proj|k_proj|v_proj weights of a block form one [3d, d] matrix (cross-attention: k|v form [2d, d]),
    the eight word_emb tables form one [1280, 256] table, the eight LM heads one [1280, d] matrix."""

    def __init__(self, d, enc_layers, dec_layers, ffn, max_pos, with_heads):
        self.d, self.enc_layers, self.dec_layers, self.ffn, self.max_pos = d, enc_layers, dec_layers, ffn, max_pos
        self.entries = {}   # name -> (offset, shape)
        self.fused = {}     # fused name -> (offset, shape)
        self.ranges = {}    # group -> (lo, hi) flat range whose gradients become final together
        self.size = 0
        self._build(with_heads)

    def _add(self, name, shape, pad=True):
        n = 1
        for s in shape:
            n *= s
        off = self.size
        self.entries[name] = (off, tuple(shape))
        self.size += n
        if pad:  # keep every (fused) tensor 32-byte aligned in fp32 / 16-byte aligned in bf16
            self.size += (-self.size) % 8
        return off

    def _fuse(self, fname, names_shapes):
        """Members are laid out back to back (no padding) so the group is one contiguous matrix/vector."""
        off0 = self.size
        rows = 0
        for name, shape in names_shapes:
            self._add(name, shape, pad=False)
            rows += shape[0]
        self.size += (-self.size) % 8
        rest = names_shapes[0][1][1:]
        self.fused[fname] = (off0, (rows,) + tuple(rest))

    def _attn(self, pre, cross):
        d = self.d
        if not cross:
            self._fuse(pre + '.wqkv', [(pre + '.q_proj.weight', (d, d)), (pre + '.k_proj.weight', (d, d)),
                                       (pre + '.v_proj.weight', (d, d))])
            self._fuse(pre + '.bqkv', [(pre + '.q_proj.bias', (d,)), (pre + '.k_proj.bias', (d,)),
                                       (pre + '.v_proj.bias', (d,))])
        else:
            self._add(pre + '.q_proj.weight', (d, d))
            self._add(pre + '.q_proj.bias', (d,))
            self._fuse(pre + '.wkv', [(pre + '.k_proj.weight', (d, d)), (pre + '.v_proj.weight', (d, d))])
            self._fuse(pre + '.bkv', [(pre + '.k_proj.bias', (d,)), (pre + '.v_proj.bias', (d,))])
        self._add(pre + '.out_proj.weight', (d, d))
        self._add(pre + '.out_proj.bias', (d,))

    def _build(self, with_heads):
        d, F = self.d, self.ffn
        assert d % 8 == 0 and F % 8 == 0
        self._fuse('emb', [('word_emb.%d.lut.weight' % i, (n, 256)) for i, n in enumerate(N_TOKENS)])
        self.emb_end = self.size
        self._add('encoder_linear.weight', (d, 2048))
        self._add('encoder_linear.bias', (d,))
        self.ranges['front'] = (0, self.size)
        for side, nl in (('encoder', self.enc_layers), ('decoder', self.dec_layers)):
            lo = self.size
            self._add('bart.%s.embed_positions.weight' % side, (self.max_pos + 2, d))
            self._add('bart.%s.layernorm_embedding.weight' % side, (d,))
            self._add('bart.%s.layernorm_embedding.bias' % side, (d,))
            self.ranges['%s.front' % side] = (lo, self.size)
            for l in range(nl):
                pre = 'bart.%s.layers.%d' % (side, l)
                lo = self.size
                self._attn(pre + '.self_attn', False)
                self._add(pre + '.self_attn_layer_norm.weight', (d,))
                self._add(pre + '.self_attn_layer_norm.bias', (d,))
                if side == 'decoder':
                    self._attn(pre + '.encoder_attn', True)
                    self._add(pre + '.encoder_attn_layer_norm.weight', (d,))
                    self._add(pre + '.encoder_attn_layer_norm.bias', (d,))
                self._add(pre + '.fc1.weight', (F, d))
                self._add(pre + '.fc1.bias', (F,))
                self._add(pre + '.fc2.weight', (d, F))
                self._add(pre + '.fc2.bias', (d,))
                self._add(pre + '.final_layer_norm.weight', (d,))
                self._add(pre + '.final_layer_norm.bias', (d,))
                self.ranges['%s.layers.%d' % (side, l)] = (lo, self.size)
        self.backbone_size = self.size
        if with_heads:
            self._fuse('heads.w', [('mask_lm.proj.%d.weight' % i, (n, d)) for i, n in enumerate(N_TOKENS)])
            self._fuse('heads.b', [('mask_lm.proj.%d.bias' % i, (n,)) for i, n in enumerate(N_TOKENS)])
            self.ranges['heads'] = (self.backbone_size, self.size)

    def off(self, name):
        if name in self.entries:
            return self.entries[name][0]
        return self.fused[name][0]


_NUM_SMS = None


def _num_sms():
    global _NUM_SMS
    if _NUM_SMS is None:
        _NUM_SMS = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count if torch.cuda.is_available() else 148
    return _NUM_SMS


def choose_split_k(n_out, n_in, M, bn, pairs, num_sms):
    """Split-K factor of a weight-gradient GEMM dW[n_out, n_in] = dY[M, n_out]^T X[M, n_in] from a wave model of the
    persistent kernel: cost(s) = ceil(tiles * s / slots) * (k-blocks / s + c), c = the per-unit pipeline fill + fp32-reduction
    epilogue in k-block units (fitted on tools/gpu_wgrad_split.py: 31 vs 39 us for 1024 x 1024, 55 vs 62 us for 2048 x 1024
    against the earlier "two units per SM" rule).  pairs: the library will run cta_group::2 tiles (256 rows, one per SM pair)."""
    tile_m = 256 if pairs else 128
    tiles = ((n_out + tile_m - 1) // tile_m) * ((n_in + bn - 1) // bn)
    slots = max(1, num_sms // 2 if pairs else num_sms)
    kblocks = (M + 63) // 64
    split, best = 1, None
    for sk in range(1, min(32, max(1, kblocks // 4)) + 1):
        cost = -(-tiles * sk // slots) * (kblocks / sk + 8.0)
        if best is None or cost < best - 1e-9:
            split, best = sk, cost
    return split


class Plan:
    """A recorded list of (function, args) library calls; run() replays it on a stream."""

    def __init__(self, dtype):
        self.lib = L.lib()
        self.dtype = dtype
        self.ops = []
        self._keep = []  # keep ctypes structs / arrays alive
        self._side_events = []
        self._side_names = set()   # GEMM launches that may run on the side stream (weight gradients)

    def _add(self, name, fn, *args):
        self.ops.append((name, fn, args))

    def join_side(self):
        """The main stream waits here for everything issued on the side stream so far (see run())."""
        self.ops.append(('join', None, ()))

    def marker(self, *payload):
        """Host-side marker (no launch): run() hands the payload to `on_marker` when it reaches it."""
        self.ops.append(('marker', None, payload))

    def run(self, stream=None, on_marker=None, profile=None, side_stream=None, on_op=None, norm_partials=None):
        """profile: optional list; every GEMM launch is then bracketed by CUDA events on the launching
        stream and (name, flops, start_event, end_event) is appended (bench.py roofline).
        side_stream: a second CUDA stream for work that has no consumer before the optimizer - the bias-gradient column sums
        (HBM-bound), D = rowsum(dO*O) of the attention backward, and the weight-gradient GEMMs recorded with side=True.  Each
        starts when its producer has finished (event) and runs next to the critical chain (dgrad GEMMs, LayerNorm and
        attention backward): the block scheduler fills the ramp and tail of every kernel with CTAs of the other stream.
        The main stream joins the side stream at every marker (the data-parallel reducer ships the layer's gradients there),
        at explicit join_side() points - recorded before every op that rewrites a buffer a side op reads - and at the end.
        norm_partials: the squared-norm partials recorded behind every 'grads_final' marker (ops 'gnorm', 'gnorm_zero'; they
        give the clip norm of the fused optimizer without a separate pass over the gradient buffer) - None: skipped,
        'local': launched (side stream), 'zero_only': only the accumulator is cleared (data parallel: the reducer adds
        the partial of each bucket behind its all-reduce)."""
        main = torch.cuda.current_stream() if stream is None else None
        s = C.c_void_p(main.cuda_stream if stream is None else stream)
        use_side = side_stream is not None and main is not None and profile is None
        ss = C.c_void_p(side_stream.cuda_stream) if use_side else None
        n = 0
        n_side = 0
        pending = False
        gemm_fns = (self.lib.pb_gemm_bf16, self.lib.pb_gemm_f32)
        for name, fn, args in self.ops:
            if fn is None:
                if pending:
                    main.wait_stream(side_stream)
                    pending = False
                if name == 'marker' and on_marker is not None:
                    on_marker(*args)
                continue
            if name == 'gnorm' and norm_partials != 'local':
                continue
            if name == 'gnorm_zero' and norm_partials is None:
                continue
            if on_op is not None and name == 'attn_bwd':
                on_op(name)     # (data-parallel reducer: launch the armed gradient buckets next to this attention backward)
            if use_side and (name in ('colsum', 'attn_prep', 'gnorm') or name in self._side_names):
                ev = self._side_events[n_side] if n_side < len(self._side_events) else None
                if ev is None:
                    ev = torch.cuda.Event()
                    self._side_events.append(ev)
                n_side += 1
                ev.record(main)
                side_stream.wait_event(ev)
                rc = fn(*args, ss)
                pending = True
            elif profile is not None and fn in gemm_fns:
                d = args[0]._obj
                flops = 2.0 * d.M * d.N * d.K * max(1, d.batch_h) * max(1, d.batch_b) * (0.5 if d.causal else 1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*args, s)
                e1.record()
                profile.append((name, flops, e0, e1))
            else:
                rc = fn(*args, s)
            n += 1
            if rc != 0:
                raise L.PBError('%s failed (%d): %s' % (name, rc, self.lib.pb_last_error().decode()))
        if pending:
            main.wait_stream(side_stream)
        return n

    # ---- op recorders ------------------------------------------------------------------
    def gemm(self, a, b, c, M, N, K, lda, ldb, ldc, bias=0, residual=0, ldr=0, a_mn=0, b_mn=0, flags=0, alpha=1.0,
             batch_h=1, batch_b=1, a_sh=0, a_sb=0, b_sh=0, b_sb=0, c_sh=0, c_sb=0, r_sh=0, r_sb=0, split_k=1,
             causal=0, aux=0, ldaux=0, r_row_mod=0, block_n=0, drop=None, name='gemm'):
        d = L.GemmDesc()
        d.a, d.b, d.c = a, b, c
        d.bias = bias or None
        d.residual = residual or None
        d.M, d.N, d.K = M, N, K
        d.a_mn_major, d.b_mn_major = a_mn, b_mn
        d.lda, d.ldb, d.ldc, d.ldr = lda, ldb, ldc, ldr
        d.batch_h, d.batch_b = batch_h, batch_b
        d.a_stride_h, d.a_stride_b, d.b_stride_h, d.b_stride_b = a_sh, a_sb, b_sh, b_sb
        d.c_stride_h, d.c_stride_b, d.r_stride_h, d.r_stride_b = c_sh, c_sb, r_sh, r_sb
        d.alpha, d.flags, d.split_k, d.causal, d.block_n = alpha, flags, split_k, causal, block_n
        d.aux = aux or None
        d.ldaux = ldaux
        d.r_row_mod = r_row_mod
        if drop is not None:
            d.drop_seed, d.drop_op, d.drop_thresh, d.drop_scale = drop
        self._keep.append(d)
        fn = self.lib.pb_gemm_bf16 if self.dtype == PB_BF16 else self.lib.pb_gemm_f32
        self._add(name, fn, C.byref(d))

    def wgrad(self, dy, x, dw, n_out, n_in, M, ld_dy, ld_x, name='wgrad', side=False):
        """dW[n_out, n_in] += dY[M, n_out]^T X[M, n_in]   (fp32 atomic accumulate, split-K)."""
        bn = int(os.environ.get('PIANOBART_B200_WGRAD_BN', '256'))
        pairs = n_out >= 1024 and bn == 256 and os.environ.get('PIANOBART_B200_CG2', '1') != '0'   # library's cta_group::2 rule
        split = choose_split_k(n_out, n_in, M, bn, pairs, _num_sms())
        self.gemm(dy, x, dw, n_out, n_in, M, ld_dy, ld_x, n_in, a_mn=1, b_mn=1,
                  flags=L.PB_GEMM_OUT_F32 | L.PB_GEMM_ATOMIC_ACC, split_k=split, block_n=bn, name=name)
        # Measured (profiles/r2_summary.md section 8): running the weight-gradient GEMMs next to the critical chain is SLOWER
        # (29.5 vs 29.1 ms per step) - two persistent tcgen05 kernels interleaving their CTAs lose more in L2 / shared-memory
        # locality than the filled ramps and tails give back under the power cap - so it stays an opt-in experiment.
        if side and os.environ.get('PIANOBART_B200_SIDE_WGRAD', '0') == '1':
            self._side_names.add(name)

    def embed_fwd(self, ids, table, out, M, ntok_arr, err=0):
        self._add('embed_fwd', self.lib.pb_octuple_embed_fwd, C.c_void_p(ids), 0, C.c_void_p(table), C.c_void_p(out),
                  C.c_longlong(M), ntok_arr, self.dtype, C.c_void_p(err or None))

    def front_fwd(self, ids, table_proj, bias, pos_rows, S, gamma, beta, y0, h0, mean, rstd, M, d, ntok_arr, drop=None, err=0):
        """Fused Octuple front end: gather-sum of the projected tables + bias + positions + LayerNorm (+ dropout)."""
        self._add('front_fwd', self.lib.pb_octuple_front_fwd, C.c_void_p(ids), 0, C.c_void_p(table_proj), C.c_void_p(bias),
                  C.c_void_p(pos_rows), S, C.c_void_p(gamma), C.c_void_p(beta), C.c_void_p(y0), C.c_void_p(h0),
                  C.c_void_p(mean), C.c_void_p(rstd), C.c_longlong(M), d, ntok_arr, C.c_float(1e-5), self._site(drop),
                  self.dtype, C.c_void_p(err or None))

    def blockdiag(self, emb, out, ntok_arr):
        self._add('front.expand', self.lib.pb_octuple_blockdiag, C.c_void_p(emb), C.c_void_p(out), 256, ntok_arr, self.dtype)

    def blockdiag_grad(self, dfull, g_emb, ntok_arr, alpha):
        self._add('front.dE_diag', self.lib.pb_octuple_blockdiag_grad, C.c_void_p(dfull), C.c_void_p(g_emb), 256, ntok_arr,
                  C.c_float(alpha))

    def onehot(self, ids, out, M, ntok_arr):
        self._add('onehot', self.lib.pb_octuple_onehot, C.c_void_p(ids), 0, C.c_void_p(out), C.c_longlong(M), ntok_arr, self.dtype)

    def grads_final(self, lo, hi, grad_base, gnorm):
        """Gradients of flat[lo:hi] are final here: marker for the data-parallel reducer + (optional, see run()) the partial
        of the clip norm over that range (norm partials fused with the gradient buckets)."""
        self.marker('grads_final', lo, hi)
        if gnorm and hi > lo:
            self._add('gnorm', self.lib.pb_sumsq, C.c_void_p(grad_base + lo * 4), C.c_longlong(hi - lo), C.c_void_p(gnorm))

    def fill_zero(self, ptr, nbytes):
        self._add('fill_zero', self.lib.pb_fill_zero, C.c_void_p(ptr), C.c_longlong(nbytes))

    def cast_from_f32(self, src, dst, n, scale=1.0):
        self._add('cast', self.lib.pb_cast_from_f32, C.c_void_p(src), C.c_void_p(dst), C.c_longlong(n), C.c_float(scale), self.dtype)

    def embed_bwd(self, ids, dx, dtable, M, ntok_arr, scale):
        self._add('embed_bwd', self.lib.pb_octuple_embed_bwd, C.c_void_p(ids), 0, C.c_void_p(dx), C.c_void_p(dtable),
                  C.c_longlong(M), ntok_arr, C.c_float(scale), self.dtype)

    def _site(self, drop):
        if drop is None:
            return None
        st = L.DropSite()
        st.seed, st.op, st.thresh, st.scale = drop
        self._keep.append(st)
        return C.byref(st)

    def ln_fwd(self, x, gamma, beta, y, mean, rstd, M, d, drop=None):
        self._add('ln_fwd', self.lib.pb_layernorm_fwd_drop, C.c_void_p(x), C.c_void_p(gamma), C.c_void_p(beta),
                  C.c_void_p(y), C.c_void_p(mean), C.c_void_p(rstd), C.c_longlong(M), d, C.c_float(1e-5),
                  self._site(drop), self.dtype)

    def ln_bwd(self, dy, x, gamma, mean, rstd, dx, dgamma, dbeta, M, d, dbias=0, dx_drop=0, in_drop=None, out_drop=None):
        self._add('ln_bwd', self.lib.pb_layernorm_bwd_drop, C.c_void_p(dy), C.c_void_p(x), C.c_void_p(gamma),
                  C.c_void_p(mean), C.c_void_p(rstd), C.c_void_p(dx), C.c_void_p(dx_drop or None), C.c_void_p(dgamma),
                  C.c_void_p(dbeta), C.c_void_p(dbias or None), C.c_longlong(M), d, self._site(in_drop),
                  self._site(out_drop), self.dtype)

    def add_u64(self, ptr, inc):
        self._add('add_u64', self.lib.pb_add_u64, C.c_void_p(ptr), C.c_ulonglong(inc))

    def softmax_fwd(self, s, p, keep, B, H, Sq, Sk, causal):
        self._add('softmax_fwd', self.lib.pb_softmax_fwd, C.c_void_p(s), C.c_void_p(p), C.c_void_p(keep or None), B, H,
                  Sq, Sk, causal, self.dtype)

    def softmax_bwd(self, p, dp, ds, keep, B, H, Sq, Sk, causal):
        self._add('softmax_bwd', self.lib.pb_softmax_bwd, C.c_void_p(p), C.c_void_p(dp), C.c_void_p(ds),
                  C.c_void_p(keep or None), B, H, Sq, Sk, causal, self.dtype)

    def attn(self, backward, q, k, v, o, ldq, ldk, ldv, ldo, lse, dvec, keep, B, H, Sq, Sk, hd, causal, dout=0, dq=0,
             dk=0, dv=0, lddq=0, lddk=0, lddv=0):
        a = L.AttnDesc()
        a.q, a.k, a.v, a.o, a.dout = q, k, v, o, dout or None
        a.dq, a.dk, a.dv = dq or None, dk or None, dv or None
        a.ldq, a.ldk, a.ldv, a.ldo, a.lddo = ldq, ldk, ldv, ldo, ldo
        a.lddq, a.lddk, a.lddv = lddq, lddk, lddv
        a.lse, a.dvec, a.key_keep = lse, dvec or None, keep or None
        a.B, a.H, a.Sq, a.Sk, a.hd, a.causal, a.scale = B, H, Sq, Sk, hd, causal, hd ** -0.5
        self._keep.append(a)
        if backward == 'prep':
            # ('attn_prep' ops go to the side stream in run(); PIANOBART_B200_SIDE_PREP=0 keeps D on the main stream)
            nm = 'attn_prep' if os.environ.get('PIANOBART_B200_SIDE_PREP', '1') != '0' else 'attn_prep_main'
            self._add(nm, self.lib.pb_attn_bwd_prep, C.byref(a))
        elif backward == 'main':
            self._add('attn_bwd', self.lib.pb_attn_bwd_main, C.byref(a))
        elif backward:
            self._add('attn_bwd', self.lib.pb_attn_bwd, C.byref(a))
        else:
            self._add('attn_fwd', self.lib.pb_attn_fwd, C.byref(a))

    def colsum(self, x, out, M, N, ld):
        self._add('colsum', self.lib.pb_colsum, C.c_void_p(x), C.c_void_p(out), C.c_longlong(M), N, C.c_longlong(ld),
                  self.dtype)


class Workspace:
    """Named device buffers for one shape; torch only allocates the memory."""

    def __init__(self, device):
        self.device = device
        self.t = {}

    def get(self, name, shape, dtype):
        key = name
        t = self.t.get(key)
        n = 1
        for s in shape:
            n *= s
        if t is None:
            t = torch.empty(max(n, 1), device=self.device, dtype=dtype)
            self.t[key] = t
        elif t.numel() < n or t.dtype != dtype:
            # recorded plans hold raw pointers: a buffer must never be reallocated
            raise RuntimeError('workspace buffer %s requested with a larger size/dtype than first allocated' % name)
        return t

    def bytes(self):
        return sum(t.numel() * t.element_size() for t in self.t.values())


class BackboneGraph:
    """Forward/backward launch plans of the PianoBART backbone (+ optional fused LM heads / CE)
    for fixed shapes.  All tensors are addressed by raw device pointers."""

    def __init__(self, layout, heads, dtype, device, B, S_enc, S_dec, w_act, w_f32, g_f32, with_heads,
                 need_backward=True, dec_embed=None, drop_p=0.0, drop_seed=None):
        """w_act: working weights (activation dtype, emb region pre-scaled by 16); w_f32: fp32 master
        (biases / LayerNorm parameters are read from it); g_f32: flat fp32 gradient buffer."""
        self.lay, self.H, self.dtype, self.device = layout, heads, dtype, device
        self.B, self.Se, self.Sd = B, S_enc, S_dec
        self.d, self.F = layout.d, layout.ffn
        self.hd = self.d // heads
        assert self.d % heads == 0
        if dtype == PB_BF16:
            assert self.hd % 8 == 0, 'bf16 mode needs head_dim % 8 == 0 (TMA 16-byte strides)'
        # fused tcgen05 attention (no S x S tensor in HBM) for the production configuration; otherwise the
        # unfused GEMM + softmax path (fp32 parity mode, head_dim != 128)
        self.flash = (dtype == PB_BF16 and self.hd == 128 and os.environ.get('PIANOBART_B200_UNFUSED_ATTN', '0') != '1')
        self.tdt = torch.bfloat16 if dtype == PB_BF16 else torch.float32
        self.es = 2 if dtype == PB_BF16 else 4
        self.w_act, self.w_f32, self.g_f32 = w_act, w_f32, g_f32
        self.ws = Workspace(device)
        self.with_heads = with_heads
        self.has_dec = S_dec > 0
        self.ntok_arr = (C.c_int * 8)(*N_TOKENS)
        self.dec_embed = dec_embed
        # training-mode dropout (HF BartConfig.dropout): masks are regenerated from (seed, site, index)
        self.drop_p = float(drop_p)
        self.drop_seed = drop_seed
        if self.drop_p > 0.0:
            assert drop_seed is not None and 0.0 < self.drop_p < 1.0
        # inputs (filled by the caller before run)
        self.enc_ids = torch.zeros(B * S_enc * 8, device=device, dtype=torch.int32)
        self.enc_keep = torch.ones(B * S_enc, device=device, dtype=torch.uint8)
        if self.has_dec:
            self.dec_ids = torch.zeros(B * S_dec * 8, device=device, dtype=torch.int32)
            self.dec_keep = torch.ones(B * S_dec, device=device, dtype=torch.uint8)
        self.err_flag = torch.zeros(1, device=device, dtype=torch.int32)
        self.fwd = Plan(dtype)
        self.bwd = Plan(dtype) if need_backward else None
        self._bwd_chunks = []
        self._build()

    def site(self, side, layer, which):
        """Dropout site descriptor (seed ptr, op id, threshold, scale) or None when dropout is off."""
        if self.drop_p <= 0.0:
            return None
        op = (0 if side == 'encoder' else 1) * 1000 + (layer + 1) * 10 + which
        thresh = min(int((1.0 - self.drop_p) * 4294967296.0), 4294967295)
        return (self.drop_seed.data_ptr(), op, thresh, 1.0 / (1.0 - self.drop_p))

    # pointer helpers -----------------------------------------------------------------------
    def W(self, name):   # working-dtype weight pointer
        return _ptr(self.w_act, self.lay.off(name))

    def Pf(self, name):  # fp32 master pointer (bias / LN params)
        return _ptr(self.w_f32, self.lay.off(name))

    def G(self, name):   # fp32 gradient pointer
        return _ptr(self.g_f32, self.lay.off(name))

    def buf(self, name, *shape, dtype=None):
        return self.ws.get(name, shape, dtype or self.tdt)

    # graph construction ----------------------------------------------------------------------
    def _build(self):
        B, d = self.B, self.d
        f, bw = self.fwd, self.bwd
        back = []  # list of closures recording backward ops; executed in reverse order at the end

        if self.drop_p > 0.0:
            f.add_u64(self.drop_seed.data_ptr(), 1)   # new masks every forward; backward re-reads the same seed
        # ---- fused front end (csrc/front.cu): T[off_a + r] = 16 E_a[r] W_a^T, tabulated once per forward for both streams.
        # The tables are first copied into their block-diagonal form Ebd [1280, 2048] (zero off-diagonal blocks, cleared once
        # here) so that T = Ebd W_in^T is ONE GEMM instead of eight latency-bound sub-GFLOP launches; the zero blocks add exact
        # zeros to the accumulators, T is bit-identical to the per-attribute products.
        es = self.es
        self.Tproj = self.buf('front.T', VOCAB, d)
        self.Ebd = self.buf('front.Ebd', VOCAB, 2048)
        self.Ebd.zero_()
        if self.Ebd.is_cuda:
            torch.cuda.current_stream().synchronize()   # plans may replay on another stream
        f.blockdiag(self.W('emb'), _ptr(self.Ebd), self.ntok_arr)
        f.gemm(_ptr(self.Ebd), self.W('encoder_linear.weight'), _ptr(self.Tproj), VOCAB, d, 2048, 2048, 2048, d, name='front.T')
        if bw is not None:
            # G = sum over tokens of onehot(m)^T dY0[m]  ([1280, d] fp32, both streams accumulate into it)
            self.Gacc = self.buf('front.G', VOCAB, d, dtype=torch.float32)
            bw.fill_zero(_ptr(self.Gacc), VOCAB * d * 4)
            # squared gradient norm accumulated range by range behind the 'grads_final' markers (Plan.run(norm_partials=...))
            self.gnorm = self.buf('gnorm', 8, dtype=torch.float32)
            self.G_base = self.g_f32.data_ptr()
            bw._add('gnorm_zero', bw.lib.pb_fill_zero, C.c_void_p(_ptr(self.gnorm)), C.c_longlong(4))
        # ---- encoder stream
        Me = B * self.Se
        enc_out, enc_back = self._stream('encoder', self.enc_ids, self.enc_keep, self.Se, None, None, 0)
        self.enc_out = enc_out
        if self.has_dec:
            dec_out, dec_back = self._stream('decoder', self.dec_ids, self.dec_keep, self.Sd, enc_out, self.enc_keep,
                                             self.Se)
            self.out = dec_out
        else:
            self.out = enc_out
        Mo = B * (self.Sd if self.has_dec else self.Se)
        self.Mo = Mo
        self.d_out = self.buf('d_out', Mo, d)  # gradient wrt last hidden state (input of backward)

        # (the heads GEMM lives in its own lazily built plan, heads_plan(): the fused heads + cross-entropy kernel of the
        # training step never materialises the fp32 logits, only PianoBartLM.forward() / the fp32 mode do)
        self._logits = None
        self._heads_plan = None

        if bw is None:
            return
        # ---- backward: heads -> decoder -> encoder
        if self.with_heads:
            self.dlogits = self.buf('dlogits', Mo, VOCAB)
            bw.colsum(_ptr(self.dlogits), self.G('heads.b'), Mo, VOCAB, VOCAB)
            bw.wgrad(_ptr(self.dlogits), _ptr(self.out), self.G('heads.w'), VOCAB, d, Mo, VOCAB, d, name='dW_heads', side=True)
            bw.gemm(_ptr(self.dlogits), self.W('heads.w'), _ptr(self.d_out), Mo, d, VOCAB, VOCAB, d, d, b_mn=1,
                    name='dH_heads')
            bw.grads_final(*self.lay.ranges['heads'], self.G_base, _ptr(self.gnorm))
        if self.has_dec:
            d_enc_out = self.buf('d_enc_out', Me, d)
            dec_back(self.d_out, d_enc_out)
            enc_back(d_enc_out, None)
        else:
            enc_back(self.d_out, None)
        # ---- front-end parameter gradients from G:  dW_in += G^T Ebd  (Ebd holds 16 E_a on its diagonal blocks) is one
        # weight-gradient GEMM straight into in_linear's [d, 2048] gradient;  dE_a = 16 G_a W_a  is the diagonal of one product
        # G W_in [1280, 2048] (8x the needed flops, 5 GFLOP) folded into the table gradients by a small kernel: 3 launches
        # instead of 16 (each 17-23 us of launch / pipeline-fill latency for < 1 GFLOP)
        Gb = self.buf('front.Gb', VOCAB, d)
        bw.cast_from_f32(_ptr(self.Gacc), _ptr(Gb), VOCAB * d)
        bw.wgrad(_ptr(Gb), _ptr(self.Ebd), self.G('encoder_linear.weight'), d, 2048, VOCAB, d, 2048, name='front.dW')
        dEfull = self.buf('front.dEfull', VOCAB, 2048, dtype=torch.float32)
        bw.gemm(_ptr(Gb), self.W('encoder_linear.weight'), _ptr(dEfull), VOCAB, 2048, d, d, 2048, 2048, b_mn=1,
                flags=L.PB_GEMM_OUT_F32, name='front.dE')
        bw.blockdiag_grad(_ptr(dEfull), self.G('emb'), self.ntok_arr, 16.0)
        bw.grads_final(*self.lay.ranges['front'], self.G_base, _ptr(self.gnorm))

    def _stream(self, side, ids, keep, S, enc_out, enc_keep, S_enc):
        """Records forward ops of one stack (front end + layers); returns (output tensor,
        backward recorder(d_out, d_enc_out))."""
        B, d, F, H, hd = self.B, self.d, self.F, self.H, self.hd
        f = self.fwd
        M = B * S
        pre = 'bart.%s' % side
        nl = self.lay.enc_layers if side == 'encoder' else self.lay.dec_layers
        is_dec = side == 'decoder'
        scale = hd ** -0.5
        T = self.tdt
        OUT32 = L.PB_GEMM_OUT_F32
        nm = lambda s: '%s.%s' % (side, s)

        custom_dec = is_dec and bool(self.dec_embed)
        # ---- front end (PianoBart.py:60-71) + positions + layernorm_embedding
        Y0 = self.buf(nm('Y0'), M, d)
        H0 = self.buf(nm('H0'), M, d)
        st0 = self.buf(nm('st0'), 2, M, dtype=torch.float32)
        if not custom_dec:
            # gathers + concat + in_linear + positions + layernorm_embedding (+ dropout) in ONE kernel; X [M,2048] never exists
            f.front_fwd(_ptr(ids), _ptr(self.Tproj), self.Pf('encoder_linear.bias'),
                        self.W(pre + '.embed_positions.weight') + 2 * d * self.es, S,
                        self.Pf(pre + '.layernorm_embedding.weight'), self.Pf(pre + '.layernorm_embedding.bias'),
                        _ptr(Y0), _ptr(H0), _ptr(st0), _ptr(st0, M), M, d, self.ntok_arr, drop=self.site(side, -1, 0),
                        err=_ptr(self.err_flag))
        else:
            # decoder input embeddings come from the caller (PianoBart.change_decoder_embedding path)
            self.dec_in = self.buf('decoder.ext_in', M, d)
            f._add('add_pos', f.lib.pb_add_rows_mod, C.c_void_p(_ptr(self.dec_in)),
                   C.c_void_p(self.W(pre + '.embed_positions.weight') + 2 * d * self.es), C.c_void_p(_ptr(Y0)),
                   C.c_longlong(M), d, S, self.dtype)
            f.ln_fwd(_ptr(Y0), self.Pf(pre + '.layernorm_embedding.weight'), self.Pf(pre + '.layernorm_embedding.bias'),
                     _ptr(H0), _ptr(st0), _ptr(st0, M), M, d, drop=self.site(side, -1, 0))

        Smax = max(self.Se, self.Sd)
        Mmax = B * Smax
        scores = None if self.flash else self.buf('scores', B * H * Smax * Smax, dtype=torch.float32)  # fp32 scratch
        layers = []
        h_in = H0
        for l in range(nl):
            lp = '%s.layers.%d' % (pre, l)
            ln = lambda s: '%s.L%d.%s' % (side, l, s)
            rec = {}
            # -- self attention
            QKV = self.buf(ln('QKV'), M, 3 * d)
            Pm = None if self.flash else self.buf(ln('P'), B * H * S * S)
            lse = self.buf(ln('lse'), B * H * S, dtype=torch.float32) if self.flash else None
            O = self.buf(ln('O'), M, d)
            A = self.buf(ln('A'), M, d)
            H1 = self.buf(ln('H1'), M, d)
            st1 = self.buf(ln('st1'), 2, M, dtype=torch.float32)
            f.gemm(_ptr(h_in), self.W(lp + '.self_attn.wqkv'), _ptr(QKV), M, 3 * d, d, d, d, 3 * d,
                   bias=self.Pf(lp + '.self_attn.bqkv'), name=ln('qkv'))
            if self.flash:
                f.attn(False, _ptr(QKV), _ptr(QKV, d), _ptr(QKV, 2 * d), _ptr(O), 3 * d, 3 * d, 3 * d, d, _ptr(lse), 0,
                       _ptr(keep), B, H, S, S, hd, 1 if is_dec else 0)
            else:
                f.gemm(_ptr(QKV), _ptr(QKV, d), _ptr(scores), S, S, hd, 3 * d, 3 * d, S, flags=OUT32, alpha=scale,
                       batch_h=H, batch_b=B, a_sh=hd, a_sb=S * 3 * d, b_sh=hd, b_sb=S * 3 * d, c_sh=S * S, c_sb=H * S * S,
                       causal=1 if is_dec else 0, name=ln('qk'))
                f.softmax_fwd(_ptr(scores), _ptr(Pm), _ptr(keep), B, H, S, S, 1 if is_dec else 0)
                f.gemm(_ptr(Pm), _ptr(QKV, 2 * d), _ptr(O), S, hd, S, S, 3 * d, d, b_mn=1, batch_h=H, batch_b=B,
                       a_sh=S * S, a_sb=H * S * S, b_sh=hd, b_sb=S * 3 * d, c_sh=hd, c_sb=S * d,
                       causal=2 if is_dec else 0, name=ln('pv'))
            f.gemm(_ptr(O), self.W(lp + '.self_attn.out_proj.weight'), _ptr(A), M, d, d, d, d, d,
                   bias=self.Pf(lp + '.self_attn.out_proj.bias'), residual=_ptr(h_in), ldr=d,
                   drop=self.site(side, l, 1), name=ln('out_proj'))
            f.ln_fwd(_ptr(A), self.Pf(lp + '.self_attn_layer_norm.weight'), self.Pf(lp + '.self_attn_layer_norm.bias'),
                     _ptr(H1), _ptr(st1), _ptr(st1, M), M, d)
            rec.update(h_in=h_in, QKV=QKV, P=Pm, lse=lse, O=O, A=A, H1=H1, st1=st1)
            h_mid = H1
            if is_dec:
                Me = B * S_enc
                Qc = self.buf(ln('Qc'), M, d)
                KVc = self.buf(ln('KVc'), Me, 2 * d)
                Pc = None if self.flash else self.buf(ln('Pc'), B * H * S * S_enc)
                lse_c = self.buf(ln('lse_c'), B * H * S, dtype=torch.float32) if self.flash else None
                Oc = self.buf(ln('Oc'), M, d)
                Ac = self.buf(ln('Ac'), M, d)
                Hc = self.buf(ln('Hc'), M, d)
                stc = self.buf(ln('stc'), 2, M, dtype=torch.float32)
                ca = lp + '.encoder_attn'
                f.gemm(_ptr(H1), self.W(ca + '.q_proj.weight'), _ptr(Qc), M, d, d, d, d, d,
                       bias=self.Pf(ca + '.q_proj.bias'), name=ln('q_c'))
                f.gemm(_ptr(enc_out), self.W(ca + '.wkv'), _ptr(KVc), Me, 2 * d, d, d, d, 2 * d,
                       bias=self.Pf(ca + '.bkv'), name=ln('kv_c'))
                if self.flash:
                    f.attn(False, _ptr(Qc), _ptr(KVc), _ptr(KVc, d), _ptr(Oc), d, 2 * d, 2 * d, d, _ptr(lse_c), 0,
                           _ptr(enc_keep), B, H, S, S_enc, hd, 0)
                else:
                    f.gemm(_ptr(Qc), _ptr(KVc), _ptr(scores), S, S_enc, hd, d, 2 * d, S_enc, flags=OUT32, alpha=scale,
                           batch_h=H, batch_b=B, a_sh=hd, a_sb=S * d, b_sh=hd, b_sb=S_enc * 2 * d, c_sh=S * S_enc,
                           c_sb=H * S * S_enc, name=ln('qk_c'))
                    f.softmax_fwd(_ptr(scores), _ptr(Pc), _ptr(enc_keep), B, H, S, S_enc, 0)
                    f.gemm(_ptr(Pc), _ptr(KVc, d), _ptr(Oc), S, hd, S_enc, S_enc, 2 * d, d, b_mn=1, batch_h=H, batch_b=B,
                           a_sh=S * S_enc, a_sb=H * S * S_enc, b_sh=hd, b_sb=S_enc * 2 * d, c_sh=hd, c_sb=S * d,
                           name=ln('pv_c'))
                f.gemm(_ptr(Oc), self.W(ca + '.out_proj.weight'), _ptr(Ac), M, d, d, d, d, d,
                       bias=self.Pf(ca + '.out_proj.bias'), residual=_ptr(H1), ldr=d, drop=self.site(side, l, 2),
                       name=ln('out_proj_c'))
                f.ln_fwd(_ptr(Ac), self.Pf(lp + '.encoder_attn_layer_norm.weight'),
                         self.Pf(lp + '.encoder_attn_layer_norm.bias'), _ptr(Hc), _ptr(stc), _ptr(stc, M), M, d)
                rec.update(Qc=Qc, KVc=KVc, Pc=Pc, lse_c=lse_c, Oc=Oc, Ac=Ac, Hc=Hc, stc=stc)
                h_mid = Hc
            # -- feed forward
            Z = self.buf(ln('Z'), M, F)
            Gt = self.buf(ln('G'), M, F)
            A2 = self.buf(ln('A2'), M, d)
            Hn = self.buf(ln('Hn'), M, d)
            st2 = self.buf(ln('st2'), 2, M, dtype=torch.float32)
            f.gemm(_ptr(h_mid), self.W(lp + '.fc1.weight'), _ptr(Gt), M, F, d, d, d, F, bias=self.Pf(lp + '.fc1.bias'),
                   flags=L.PB_GEMM_GELU | L.PB_GEMM_AUX_DGELU, aux=_ptr(Z), ldaux=F, name=ln('fc1'))
            f.gemm(_ptr(Gt), self.W(lp + '.fc2.weight'), _ptr(A2), M, d, F, F, F, d, bias=self.Pf(lp + '.fc2.bias'),
                   residual=_ptr(h_mid), ldr=d, drop=self.site(side, l, 3), name=ln('fc2'))
            f.ln_fwd(_ptr(A2), self.Pf(lp + '.final_layer_norm.weight'), self.Pf(lp + '.final_layer_norm.bias'),
                     _ptr(Hn), _ptr(st2), _ptr(st2, M), M, d)
            rec.update(h_mid=h_mid, Z=Z, G=Gt, A2=A2, Hn=Hn, st2=st2, lp=lp)
            layers.append(rec)
            h_in = Hn
        out = h_in

        def record_backward(d_out, d_enc_out):
            """d_out: gradient wrt `out` (overwritten); d_enc_out: accumulator for the encoder output grad."""
            bw = self.bwd
            # scratch gradient buffers shared by all layers of both stacks
            dA = self.buf('g.dA', Mmax, d)
            # gradient of the dropped-out sub-layer output (== dA when dropout is off)
            dAd = self.buf('g.dAd', Mmax, d) if self.drop_p > 0.0 else dA
            dZ = self.buf('g.dZ', Mmax, F)
            dH1 = self.buf('g.dH1', Mmax, d)
            dO = self.buf('g.dO', Mmax, d)
            dQKV = self.buf('g.dQKV', Mmax, 3 * d)
            dcur = d_out
            dnext = self.buf('g.dH_' + side, M, d)
            first_cross = True
            for l in reversed(range(nl)):
                r = layers[l]
                lp = r['lp']
                ln = lambda s: '%s.L%d.%s' % (side, l, s)
                # -- FFN backward
                bw.join_side()     # (the LayerNorm backward rewrites dAd, which side-stream weight gradients read)
                bw.ln_bwd(_ptr(dcur), _ptr(r['A2']), self.Pf(lp + '.final_layer_norm.weight'), _ptr(r['st2']),
                          _ptr(r['st2'], M), _ptr(dA), self.G(lp + '.final_layer_norm.weight'),
                          self.G(lp + '.final_layer_norm.bias'), M, d, dbias=self.G(lp + '.fc2.bias'),
                          dx_drop=_ptr(dAd), out_drop=self.site(side, l, 3))
                bw.wgrad(_ptr(dAd), _ptr(r['G']), self.G(lp + '.fc2.weight'), d, F, M, d, F, name=ln('dW_fc2'), side=True)
                bw.gemm(_ptr(dAd), self.W(lp + '.fc2.weight'), _ptr(dZ), M, F, d, d, F, F, b_mn=1,
                        flags=L.PB_GEMM_MUL_AUX, aux=_ptr(r['Z']), ldaux=F, name=ln('dZ'))
                bw.colsum(_ptr(dZ), self.G(lp + '.fc1.bias'), M, F, F)
                bw.wgrad(_ptr(dZ), _ptr(r['h_mid']), self.G(lp + '.fc1.weight'), F, d, M, F, d, name=ln('dW_fc1'), side=True)
                bw.gemm(_ptr(dZ), self.W(lp + '.fc1.weight'), _ptr(dH1), M, d, F, F, d, d, b_mn=1, residual=_ptr(dA),
                        ldr=d, name=ln('dH_mid'))
                dmid = dH1
                if is_dec:
                    Me = B * S_enc
                    ca = lp + '.encoder_attn'
                    dQc = self.buf('g.dQc', Mmax, d)
                    dKVc = self.buf('g.dKVc', Mmax, 2 * d)
                    bw.join_side()
                    bw.ln_bwd(_ptr(dmid), _ptr(r['Ac']), self.Pf(lp + '.encoder_attn_layer_norm.weight'),
                              _ptr(r['stc']), _ptr(r['stc'], M), _ptr(dA), self.G(lp + '.encoder_attn_layer_norm.weight'),
                              self.G(lp + '.encoder_attn_layer_norm.bias'), M, d, dbias=self.G(ca + '.out_proj.bias'),
                              dx_drop=_ptr(dAd), out_drop=self.site(side, l, 2))
                    # dP = dO V^T ; dV = P^T dO ; dS = softmax'(P, dP) ; dQ = scale dS K ; dK = scale dS^T Q
                    if self.flash:
                        # dO first, then D = rowsum(dO * O) on the side stream next to the out_proj weight gradient
                        bw.gemm(_ptr(dAd), self.W(ca + '.out_proj.weight'), _ptr(dO), M, d, d, d, d, d, b_mn=1, name=ln('dOc'))
                        dvec = self.buf('g.dvec', B * H * Smax, dtype=torch.float32)
                        aargs = (_ptr(r['Qc']), _ptr(r['KVc']), _ptr(r['KVc'], d), _ptr(r['Oc']), d, 2 * d, 2 * d, d,
                                 _ptr(r['lse_c']), _ptr(dvec), _ptr(enc_keep), B, H, S, S_enc, hd, 0)
                        akw = dict(dout=_ptr(dO), dq=_ptr(dQc), dk=_ptr(dKVc), dv=_ptr(dKVc, d), lddq=d, lddk=2 * d, lddv=2 * d)
                        bw.attn('prep', *aargs, **akw)
                        bw.wgrad(_ptr(dAd), _ptr(r['Oc']), self.G(ca + '.out_proj.weight'), d, d, M, d, d, name=ln('dW_oc'), side=True)
                        bw.join_side()
                        bw.attn('main', *aargs, **akw)
                    else:
                        bw.wgrad(_ptr(dAd), _ptr(r['Oc']), self.G(ca + '.out_proj.weight'), d, d, M, d, d, name=ln('dW_oc'), side=True)
                        bw.gemm(_ptr(dAd), self.W(ca + '.out_proj.weight'), _ptr(dO), M, d, d, d, d, d, b_mn=1, name=ln('dOc'))
                        bw.gemm(_ptr(dO), _ptr(r['KVc'], d), _ptr(scores), S, S_enc, hd, d, 2 * d, S_enc, flags=OUT32,
                                batch_h=H, batch_b=B, a_sh=hd, a_sb=S * d, b_sh=hd, b_sb=S_enc * 2 * d, c_sh=S * S_enc,
                                c_sb=H * S * S_enc, name=ln('dP_c'))
                        bw.gemm(_ptr(r['Pc']), _ptr(dO), _ptr(dKVc, d), S_enc, hd, S, S_enc, d, 2 * d, a_mn=1, b_mn=1,
                                batch_h=H, batch_b=B, a_sh=S * S_enc, a_sb=H * S * S_enc, b_sh=hd, b_sb=S * d, c_sh=hd,
                                c_sb=S_enc * 2 * d, name=ln('dV_c'))
                        bw.softmax_bwd(_ptr(r['Pc']), _ptr(scores), _ptr(r['Pc']), _ptr(enc_keep), B, H, S, S_enc, 0)
                        bw.gemm(_ptr(r['Pc']), _ptr(r['KVc']), _ptr(dQc), S, hd, S_enc, S_enc, 2 * d, d, b_mn=1,
                                alpha=scale, batch_h=H, batch_b=B, a_sh=S * S_enc, a_sb=H * S * S_enc, b_sh=hd,
                                b_sb=S_enc * 2 * d, c_sh=hd, c_sb=S * d, name=ln('dQ_c'))
                        bw.gemm(_ptr(r['Pc']), _ptr(r['Qc']), _ptr(dKVc), S_enc, hd, S, S_enc, d, 2 * d, a_mn=1, b_mn=1,
                                alpha=scale, batch_h=H, batch_b=B, a_sh=S * S_enc, a_sb=H * S * S_enc, b_sh=hd,
                                b_sb=S * d, c_sh=hd, c_sb=S_enc * 2 * d, name=ln('dK_c'))
                    bw.colsum(_ptr(dQc), self.G(ca + '.q_proj.bias'), M, d, d)
                    bw.wgrad(_ptr(dQc), _ptr(r['H1']), self.G(ca + '.q_proj.weight'), d, d, M, d, d, name=ln('dW_qc'), side=True)
                    bw.colsum(_ptr(dKVc), self.G(ca + '.bkv'), Me, 2 * d, 2 * d)
                    bw.wgrad(_ptr(dKVc), _ptr(enc_out), self.G(ca + '.wkv'), 2 * d, d, Me, 2 * d, d, name=ln('dW_kvc'), side=True)
                    bw.gemm(_ptr(dKVc), self.W(ca + '.wkv'), _ptr(d_enc_out), Me, d, 2 * d, 2 * d, d, d, b_mn=1,
                            residual=0 if first_cross else _ptr(d_enc_out), ldr=d, name=ln('dEnc'))
                    first_cross = False
                    dmid2 = self.buf('g.dH1b', Mmax, d)
                    bw.gemm(_ptr(dQc), self.W(ca + '.q_proj.weight'), _ptr(dmid2), M, d, d, d, d, d, b_mn=1,
                            residual=_ptr(dA), ldr=d, name=ln('dH1_c'))
                    dmid = dmid2
                # -- self attention backward
                sa = lp + '.self_attn'
                cz = 1 if is_dec else 0
                bw.join_side()
                bw.ln_bwd(_ptr(dmid), _ptr(r['A']), self.Pf(lp + '.self_attn_layer_norm.weight'), _ptr(r['st1']),
                          _ptr(r['st1'], M), _ptr(dA), self.G(lp + '.self_attn_layer_norm.weight'),
                          self.G(lp + '.self_attn_layer_norm.bias'), M, d, dbias=self.G(sa + '.out_proj.bias'),
                          dx_drop=_ptr(dAd), out_drop=self.site(side, l, 1))
                QKV, Pm = r['QKV'], r['P']
                if self.flash:
                    bw.gemm(_ptr(dAd), self.W(sa + '.out_proj.weight'), _ptr(dO), M, d, d, d, d, d, b_mn=1, name=ln('dO'))
                    dvec = self.buf('g.dvec', B * H * Smax, dtype=torch.float32)
                    aargs = (_ptr(QKV), _ptr(QKV, d), _ptr(QKV, 2 * d), _ptr(r['O']), 3 * d, 3 * d, 3 * d, d,
                             _ptr(r['lse']), _ptr(dvec), _ptr(keep), B, H, S, S, hd, cz)
                    akw = dict(dout=_ptr(dO), dq=_ptr(dQKV), dk=_ptr(dQKV, d), dv=_ptr(dQKV, 2 * d), lddq=3 * d, lddk=3 * d,
                               lddv=3 * d)
                    bw.attn('prep', *aargs, **akw)
                    bw.wgrad(_ptr(dAd), _ptr(r['O']), self.G(sa + '.out_proj.weight'), d, d, M, d, d, name=ln('dW_o'), side=True)
                    bw.join_side()
                    bw.attn('main', *aargs, **akw)
                else:
                    bw.wgrad(_ptr(dAd), _ptr(r['O']), self.G(sa + '.out_proj.weight'), d, d, M, d, d, name=ln('dW_o'), side=True)
                    bw.gemm(_ptr(dAd), self.W(sa + '.out_proj.weight'), _ptr(dO), M, d, d, d, d, d, b_mn=1, name=ln('dO'))
                    bw.gemm(_ptr(dO), _ptr(QKV, 2 * d), _ptr(scores), S, S, hd, d, 3 * d, S, flags=OUT32, batch_h=H,
                            batch_b=B, a_sh=hd, a_sb=S * d, b_sh=hd, b_sb=S * 3 * d, c_sh=S * S, c_sb=H * S * S, causal=cz,
                            name=ln('dP'))
                    bw.gemm(_ptr(Pm), _ptr(dO), _ptr(dQKV, 2 * d), S, hd, S, S, d, 3 * d, a_mn=1, b_mn=1, batch_h=H,
                            batch_b=B, a_sh=S * S, a_sb=H * S * S, b_sh=hd, b_sb=S * d, c_sh=hd, c_sb=S * 3 * d,
                            name=ln('dV'))
                    bw.softmax_bwd(_ptr(Pm), _ptr(scores), _ptr(Pm), _ptr(keep), B, H, S, S, cz)
                    bw.gemm(_ptr(Pm), _ptr(QKV, d), _ptr(dQKV), S, hd, S, S, 3 * d, 3 * d, b_mn=1, alpha=scale, batch_h=H,
                            batch_b=B, a_sh=S * S, a_sb=H * S * S, b_sh=hd, b_sb=S * 3 * d, c_sh=hd, c_sb=S * 3 * d,
                            causal=2 if is_dec else 0, name=ln('dQ'))
                    bw.gemm(_ptr(Pm), _ptr(QKV), _ptr(dQKV, d), S, hd, S, S, 3 * d, 3 * d, a_mn=1, b_mn=1, alpha=scale,
                            batch_h=H, batch_b=B, a_sh=S * S, a_sb=H * S * S, b_sh=hd, b_sb=S * 3 * d, c_sh=hd,
                            c_sb=S * 3 * d, name=ln('dK'))
                bw.colsum(_ptr(dQKV), self.G(sa + '.bqkv'), M, 3 * d, 3 * d)
                bw.wgrad(_ptr(dQKV), _ptr(r['h_in']), self.G(sa + '.wqkv'), 3 * d, d, M, 3 * d, d, name=ln('dW_qkv'), side=True)
                bw.gemm(_ptr(dQKV), self.W(sa + '.wqkv'), _ptr(dnext), M, d, 3 * d, 3 * d, d, d, b_mn=1,
                        residual=_ptr(dA), ldr=d, name=ln('dH_in'))
                dcur, dnext = dnext, dcur
                bw.grads_final(*self.lay.ranges['%s.layers.%d' % (side, l)], self.G_base, _ptr(self.gnorm))
            # -- front end backward
            # (the gradient wrt caller-provided decoder input embeddings outlives this stream's backward: it gets its own
            # buffer - the shared scratch dA is rewritten by the encoder stream's backward that follows)
            dY0 = self.buf('g.d_dec_in', M, d) if custom_dec else dA
            bw.join_side()
            bw.ln_bwd(_ptr(dcur), _ptr(Y0), self.Pf(pre + '.layernorm_embedding.weight'), _ptr(st0), _ptr(st0, M),
                      _ptr(dY0), self.G(pre + '.layernorm_embedding.weight'), self.G(pre + '.layernorm_embedding.bias'),
                      M, d, dbias=0 if custom_dec else self.G('encoder_linear.bias'), in_drop=self.site(side, -1, 0))
            # d pos[s+2] += sum_b dY0[b, s]  == column sums of dY0 viewed as [B, S*d]
            bw.colsum(_ptr(dY0), self.G(pre + '.embed_positions.weight') + 2 * d * 4, B, S * d, S * d)
            if not custom_dec:
                onehot = self.buf('g.onehot', Mmax, VOCAB)
                bw.onehot(_ptr(ids), _ptr(onehot), M, self.ntok_arr)
                bw.wgrad(_ptr(onehot), _ptr(dY0), _ptr(self.Gacc), VOCAB, d, M, VOCAB, d, name=nm('dG'), side=True)
            else:
                self.d_dec_in = dY0      # gradient wrt the caller's decoder input embeddings
            bw.grads_final(*self.lay.ranges['%s.front' % side], self.G_base, _ptr(self.gnorm))

        return out, record_backward

    # ---------------------------------------------------------------------------------------
    def set_inputs(self, enc_ids, enc_keep, dec_ids=None, dec_keep=None):
        """Copies (device-to-device, or host-to-device from pinned memory) the step inputs into the
        graph's static input buffers.  ids: (B,S,8) integer tensor; keep: (B,S) any dtype, non-zero = keep."""
        self.enc_ids.copy_(enc_ids.reshape(-1), non_blocking=True)
        if enc_keep is None:
            self.enc_keep.fill_(1)
        else:
            self.enc_keep.copy_((enc_keep.reshape(-1) != 0), non_blocking=True)
        if self.has_dec:
            if dec_ids is not None:
                self.dec_ids.copy_(dec_ids.reshape(-1), non_blocking=True)
            if dec_keep is None:
                self.dec_keep.fill_(1)
            else:
                self.dec_keep.copy_((dec_keep.reshape(-1) != 0), non_blocking=True)

    @property
    def logits(self):
        """fp32 logits [Mo, 1280] of the eight MLM heads (allocated on first use)"""
        if self._logits is None:
            self._logits = self.buf('logits', self.Mo, VOCAB, dtype=torch.float32)
        return self._logits

    def heads_plan(self):
        """model.py:119-126 as one N = 1280 GEMM with bias into `logits`"""
        if self._heads_plan is None:
            d = self.d
            hp = Plan(self.dtype)
            hp.gemm(_ptr(self.out), self.W('heads.w'), _ptr(self.logits), self.Mo, VOCAB, d, d, d, VOCAB,
                    bias=self.Pf('heads.b'), flags=L.PB_GEMM_OUT_F32, name='heads')
            self._heads_plan = hp
        return self._heads_plan

    def forward(self):
        n = self.fwd.run()
        if self.with_heads:
            n += self.heads_plan().run()
        return n

    def backward(self):
        return self.bwd.run(side_stream=self.side_stream())

    def side_stream(self):
        """second stream for the bias-gradient column sums of the backward plan (Plan.run); PIANOBART_B200_SIDE_COLSUM=0 off"""
        if os.environ.get('PIANOBART_B200_SIDE_COLSUM', '1') == '0':
            return None
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side


def profile_gemms(step):
    """Runs one training step of `step` (PretrainStep) with every GEMM launch timed by CUDA events.
    Returns (total algorithmic GEMM flops, total GEMM milliseconds, number of GEMM launches)."""
    prof = []
    step.run(train=True, profile=prof)
    torch.cuda.synchronize()
    # the dominant kernel = the large projections / weight-gradient products; the ~25 sub-GFLOP launches of the fused front
    # end (vocabulary-sized table products, a few microseconds each) are not part of that aggregate
    big = [p for p in prof if p[1] >= 1e9]
    flops = sum(p[1] for p in big)
    ms = sum(p[2].elapsed_time(p[3]) for p in big)
    return flops, ms, len(big)

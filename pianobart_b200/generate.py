"""KV-cache autoregressive generation, replacing the O(S^2) loop of reference
`PianoBartLM.forward(generate=True)` (model.py:28-66) and `sample`/`sampling`/`nucleus` (model.py:68-107).

Semantics kept from the reference: decoder starts from the <SOS> row, every step samples the 8 attributes
with temperatures t=[1.2,1.2,5,1,2,5,5,1.2] and nucleus p=[1,1,1,.9,.9,1,1,.9] (p == 1 degenerates to greedy),
one numpy uniform is consumed per attribute per step in attribute order (np.random.choice), generation stops at
the first step where any attribute >= its <PAD> id and that step is not written, the result is PAD-filled
int64 (B,S,8).  Batch sizes > 1 are a new capability (the reference exits unless batch == 1).

One decode step = one CUDA-graph replay of ~120 launches; the step index, the sampled token hand-over and the
stop flags live in device memory so the same graph is replayed for every position.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L
from . import engine as E

SAMPLE_T = [1.2, 1.2, 5, 1, 2, 5, 5, 1.2]      # model.py:70
SAMPLE_P = [1, 1, 1, 0.9, 0.9, 1, 1, 0.9]      # model.py:71


class Generator:
    def __init__(self, lm, B, S_enc, S_max=None, use_graph=True, force_gemm=False):
        pb = lm.pianobart
        pb._ensure_packed()
        if pb.pb_dtype != E.PB_BF16:
            raise L.PBError('KV-cache decode is implemented for the bf16 production mode only')
        self.lm, self.pb = lm, pb
        self._pack_gen = pb._pack_gen
        self.B, self.Se = B, S_enc
        self.S = S_max or S_enc
        self.use_graph = use_graph
        self.force_gemm = force_gemm     # use the tcgen05 split-K path even for tiny batches (tests)
        lay = pb.layout
        d, F, H = lay.d, lay.ffn, pb.heads
        self.d, self.F, self.H, self.hd = d, F, H, d // H
        dev = pb._flat.device
        self.dev = dev
        self.lib = L.lib()
        bf, f32, i32 = torch.bfloat16, torch.float32, torch.int32
        z = lambda *s, dt=bf: torch.zeros(*s, device=dev, dtype=dt)
        S = self.S
        self.t_dev = z(1, dt=i32)
        self.cur_tok = z(B, 8, dt=i32)
        self.result = z(B, S, 8, dt=i32)
        self.sampled = z(B, S, 8, dt=i32)
        self.done = z(B, dt=i32)
        self.n_written = z(B, dt=i32)
        self.uniforms = z(B, S, 8, dt=torch.float64)
        self.forced = None
        self.x_emb = z(B, 2048)
        nmax = max(3 * d, F, E.VOCAB, 2 * d)
        self.acc = z(B, nmax, dt=f32)
        self.hA, self.hB, self.h1, self.h2 = z(B, d), z(B, d), z(B, d), z(B, d)
        self.qkv, self.qc, self.ao, self.f1 = z(B, 3 * d), z(B, d), z(B, d), z(B, F)
        self.logits = z(B, E.VOCAB, dt=f32)
        ns = (max(S, S_enc) + 127) // 128
        self.attn_ws = z(B * H * ns * (self.hd + 2), dt=f32)
        self.attn_tickets = z(B * H, dt=i32)
        self.self_cache = None      # (allocated below only for the per-op paths)
        self.cross_kv = [z(B * S_enc, 2 * d) for _ in range(lay.dec_layers)]
        self.enc_graph = pb._graph(B, S_enc, 0, False, False, 0.0)  # no dropout (demo.py:149-150 behaviour)
        # batch 1, default geometry: one persistent cooperative kernel generates many tokens per launch
        default_geom = (d == 1024 and F == 2048 and H == 8 and 1 <= lay.dec_layers <= 8 and max(self.S, S_enc) <= 18 * 57
                        and not force_gemm and os.environ.get('PIANOBART_B200_DECODE_PERSIST', '1') != '0')
        self.persist = default_geom and B == 1
        # batch 64, default geometry: the batched persistent kernel (csrc/decode_batch.cu)
        self.persist_batch = default_geom and B == 64
        if not (self.persist or self.persist_batch):
            self.self_cache = [z(B, S, 2 * d) for _ in range(lay.dec_layers)]
        self.ntok_arr = (C.c_int * 8)(*E.N_TOKENS)
        self.pad_arr = (C.c_int * 8)(*[int(x) for x in pb.pad_word_np])
        self.temp_arr = (C.c_float * 8)(*[float(x) for x in SAMPLE_T])
        self.p_arr = (C.c_float * 8)(*[float(x) for x in SAMPLE_P])
        self.sos = torch.tensor(pb.sos_word_np, dtype=i32, device=dev)
        self.prefill = E.Plan(E.PB_BF16)
        self.step = E.Plan(E.PB_BF16)
        self._build()
        self.graph = None

    # ------------------------------------------------------------------
    def _W(self, name):
        return E._ptr(self.pb._wact, self.pb.layout.off(name))

    def _Pf(self, name):
        return E._ptr(self.pb._flat, self.pb.layout.off(name))

    def _skinny(self, plan, x, w, N, K, name):
        """acc[B, N] += x[B, K] W[N, K]^T : tcgen05 GEMM with M = batch, split over K so that every SM streams a
        slice of the weight matrix exactly once."""
        tiles = (N + 127) // 128
        kblocks = (K + 63) // 64
        split = max(1, min(kblocks, (148 + tiles - 1) // tiles))
        plan.gemm(E._ptr(x), w, E._ptr(self.acc), self.B, N, K, K, K, N, flags=L.PB_GEMM_OUT_F32 | L.PB_GEMM_ATOMIC_ACC,
                  split_k=split, name=name)
        plan.ops[-1][2][0]._obj.block_n = 128

    def _finalize(self, plan, N, bias=0, residual=None, pos=0, ln=None, out=None, out_f32=None, gelu=0):
        P = C.c_void_p
        g, b = (self._Pf(ln + '.weight'), self._Pf(ln + '.bias')) if ln else (0, 0)
        plan._add('decode_finalize', self.lib.pb_decode_finalize, P(E._ptr(self.acc)), P(bias or None),
                  P(E._ptr(residual) if residual is not None else None), P(pos or None), P(E._ptr(self.t_dev)), P(g or None),
                  P(b or None), P(E._ptr(out) if out is not None else None),
                  P(E._ptr(out_f32) if out_f32 is not None else None), self.B, N, gelu)

    def _attn(self, plan, q, q_ld, k_new, v_new, kc, vc, kv_bs, kv_ld, keep, n_keys, append, out, max_keys, out_f32=None):
        P = C.c_void_p
        plan._add('decode_attn', self.lib.pb_decode_attn, P(q), q_ld, P(k_new or None), P(v_new or None), P(kc), P(vc),
                  C.c_longlong(kv_bs), kv_ld, P(keep or None), n_keys, P(E._ptr(self.t_dev)), append,
                  P(E._ptr(out) if out is not None else None), self.d,
                  self.B, self.H, self.hd, C.c_float(self.hd ** -0.5), max_keys, P(E._ptr(self.attn_ws)),
                  P(E._ptr(self.attn_tickets)), P(E._ptr(out_f32) if out_f32 is not None else None))

    def _gemv(self, plan, x_raw, W, N, K, ln=None, x_norm_out=None, bias=0, residual=None, pos=0, y_f32=None, y_bf16=None,
              gelu=0):
        P = C.c_void_p
        g, b = (self._Pf(ln + '.weight'), self._Pf(ln + '.bias')) if ln else (0, 0)
        tp = lambda t: P(E._ptr(t) if t is not None else None)
        plan._add('decode_gemv', self.lib.pb_decode_gemv, tp(x_raw), P(g or None), P(b or None), tp(x_norm_out), P(W),
                  P(bias or None), tp(residual), P(pos or None), P(E._ptr(self.t_dev)), tp(y_f32), tp(y_bf16), self.B, N, K, gelu)

    def _build_small_batch(self):
        """B <= 8: HBM-bound GEMV path - one launch per projection, LayerNorms fused into the consumer's prologue."""
        pb, lay = self.pb, self.pb.layout
        B, d, F, Se, S = self.B, self.d, self.F, self.Se, self.S
        eg = self.enc_graph
        f32 = torch.float32
        z = lambda *s: torch.zeros(*s, device=self.dev, dtype=f32)
        self.x_emb32 = z(B, 2048)
        self.raw = [z(B, d) for _ in range(4)]
        self.hn, self.h1n, self.h2n = z(B, d), z(B, d), z(B, d)
        self.ao32, self.f1_32 = z(B, d), z(B, F)
        for l in range(lay.dec_layers):
            ca = 'bart.decoder.layers.%d.encoder_attn' % l
            self.prefill.gemm(E._ptr(eg.enc_out), self._W(ca + '.wkv'), E._ptr(self.cross_kv[l]), B * Se, 2 * d, d, d, d, 2 * d,
                              bias=self._Pf(ca + '.bkv'), name='kv_c%d' % l)
        st = self.step
        P = C.c_void_p
        st._add('embed', self.lib.pb_octuple_embed_fwd, P(E._ptr(self.cur_tok)), 0, P(self._W('emb')), P(E._ptr(self.x_emb)),
                C.c_longlong(B), self.ntok_arr, E.PB_BF16, P(None))
        st._add('cast', self.lib.pb_cast_to_f32, P(E._ptr(self.x_emb)), P(E._ptr(self.x_emb32)), C.c_longlong(B * 2048), E.PB_BF16)
        h_raw = self.raw[0]
        self._gemv(st, self.x_emb32, self._W('encoder_linear.weight'), d, 2048, bias=self._Pf('encoder_linear.bias'),
                   pos=self._W('bart.decoder.embed_positions.weight'), y_f32=h_raw)
        ln_prev = 'bart.decoder.layernorm_embedding'
        for l in range(lay.dec_layers):
            lp = 'bart.decoder.layers.%d' % l
            sa, ca = lp + '.self_attn', lp + '.encoder_attn'
            r1, r2, r3 = self.raw[1], self.raw[2], self.raw[3] if h_raw is self.raw[0] else self.raw[0]
            self._gemv(st, h_raw, self._W(sa + '.wqkv'), 3 * d, d, ln=ln_prev, x_norm_out=self.hn, bias=self._Pf(sa + '.bqkv'),
                       y_bf16=self.qkv)
            cache = self.self_cache[l]
            self._attn(st, E._ptr(self.qkv), 3 * d, E._ptr(self.qkv, d), E._ptr(self.qkv, 2 * d), E._ptr(cache), E._ptr(cache, d),
                       S * 2 * d, 2 * d, 0, 0, 1, None, S, out_f32=self.ao32)
            self._gemv(st, self.ao32, self._W(sa + '.out_proj.weight'), d, d, bias=self._Pf(sa + '.out_proj.bias'),
                       residual=self.hn, y_f32=r1)
            self._gemv(st, r1, self._W(ca + '.q_proj.weight'), d, d, ln=lp + '.self_attn_layer_norm', x_norm_out=self.h1n,
                       bias=self._Pf(ca + '.q_proj.bias'), y_bf16=self.qc)
            kv = self.cross_kv[l]
            self._attn(st, E._ptr(self.qc), d, 0, 0, E._ptr(kv), E._ptr(kv, d), Se * 2 * d, 2 * d, E._ptr(eg.enc_keep), Se, 0,
                       None, Se, out_f32=self.ao32)
            self._gemv(st, self.ao32, self._W(ca + '.out_proj.weight'), d, d, bias=self._Pf(ca + '.out_proj.bias'),
                       residual=self.h1n, y_f32=r2)
            self._gemv(st, r2, self._W(lp + '.fc1.weight'), F, d, ln=lp + '.encoder_attn_layer_norm', x_norm_out=self.h2n,
                       bias=self._Pf(lp + '.fc1.bias'), y_f32=self.f1_32, gelu=1)
            self._gemv(st, self.f1_32, self._W(lp + '.fc2.weight'), d, F, bias=self._Pf(lp + '.fc2.bias'), residual=self.h2n,
                       y_f32=r3)
            h_raw, ln_prev = r3, lp + '.final_layer_norm'
        self._gemv(st, h_raw, self._W('heads.w'), E.VOCAB, d, ln=ln_prev, bias=self._Pf('heads.b'), y_f32=self.logits)
        self._sample_idx = len(st.ops)
        st._add('sample', self.lib.pb_decode_sample, P(E._ptr(self.logits)), P(E._ptr(self.uniforms)), P(None),
                P(E._ptr(self.t_dev)), P(E._ptr(self.cur_tok)), P(E._ptr(self.sampled)), B, S, self.ntok_arr, self.temp_arr,
                self.p_arr)
        st._add('advance', self.lib.pb_decode_advance, P(E._ptr(self.cur_tok)), P(E._ptr(self.result)), P(E._ptr(self.done)),
                P(E._ptr(self.t_dev)), P(E._ptr(self.n_written)), B, S, self.pad_arr)
        self.launches_per_step = len(st.ops)

    def _build_persist(self):
        """Batch 1: csrc/decode_persist.cu.  Prefill = cross-attention K/V GEMMs + relayout into the split-major cache
        layout; the decode step itself is not a plan but one cooperative launch per run_steps() call."""
        pb, lay = self.pb, self.pb.layout
        d, F, Se, S = self.d, self.F, self.Se, self.S
        eg = self.enc_graph
        P = C.c_void_p
        dev = self.dev
        zb = lambda *s: torch.zeros(*s, device=dev, dtype=torch.bfloat16)
        zw = lambda n: torch.zeros(n, device=dev, dtype=torch.int64)
        nl = lay.dec_layers
        self.self_k_r = [zb(8, 18, 57, 128) for _ in range(nl)]
        self.self_v_r = [zb(8, 18, 57, 128) for _ in range(nl)]
        self.cross_k_r = [zb(8, 18, 57, 128) for _ in range(nl)]
        self.cross_v_r = [zb(8, 18, 57, 128) for _ in range(nl)]
        for l in range(nl):
            ca = 'bart.decoder.layers.%d.encoder_attn' % l
            self.prefill.gemm(E._ptr(eg.enc_out), self._W(ca + '.wkv'), E._ptr(self.cross_kv[l]), Se, 2 * d, d, d, d, 2 * d,
                              bias=self._Pf(ca + '.bkv'), name='kv_c%d' % l)
            self.prefill._add('kv_relayout', self.lib.pb_decode_kv_relayout, P(E._ptr(self.cross_kv[l])),
                              P(E._ptr(self.cross_k_r[l])), P(E._ptr(self.cross_v_r[l])), Se)
        self.ll = dict(raw0=zw(2 * d), raw1=zw(2 * d), raw2=zw(2 * d), qkv=zw(6 * d), qc=zw(2 * d), ob=zw(2 * d),
                       f1=zw(2 * F), part=zw(8 * 18 * 132), logits_ll=zw(E.VOCAB), tok_ll=zw(8))
        self.epoch = torch.zeros(1, device=dev, dtype=torch.int32)
        self.err_flag = torch.zeros(1, device=dev, dtype=torch.int32)
        D = L.DecodePersistDesc()
        self._fill_layers(D)
        D.n_layers, D.S_enc, D.S_max, D.stop_when_done = nl, Se, S, 0
        self._fill_common(D)
        D.enc_keep = E._ptr(eg.enc_keep)
        D.logits_out = E._ptr(self.logits)
        for k, t in self.ll.items():
            setattr(D, k, E._ptr(t))
        D.epoch, D.error_flag = E._ptr(self.epoch), E._ptr(self.err_flag)
        self.pdesc = D
        self.launches_per_step = 1
        self._steps_issued = 0

    def _build_persist_batch(self):
        """Batch 64: csrc/decode_batch.cu.  Prefill = cross-attention K/V GEMMs + relayout to [seq][head][key][128]."""
        pb, lay = self.pb, self.pb.layout
        B, d, F, Se, S = self.B, self.d, self.F, self.Se, self.S
        eg = self.enc_graph
        P = C.c_void_p
        dev = self.dev
        zb = lambda *s: torch.zeros(*s, device=dev, dtype=torch.bfloat16)
        nl = lay.dec_layers
        self.self_k_r = [zb(B, 8, S, 128) for _ in range(nl)]
        self.self_v_r = [zb(B, 8, S, 128) for _ in range(nl)]
        self.cross_k_r = [zb(B, 8, Se, 128) for _ in range(nl)]
        self.cross_v_r = [zb(B, 8, Se, 128) for _ in range(nl)]
        for l in range(nl):
            ca = 'bart.decoder.layers.%d.encoder_attn' % l
            self.prefill.gemm(E._ptr(eg.enc_out), self._W(ca + '.wkv'), E._ptr(self.cross_kv[l]), B * Se, 2 * d, d, d, d, 2 * d,
                              bias=self._Pf(ca + '.bkv'), name='kv_c%d' % l)
            self.prefill._add('kv_relayout', self.lib.pb_decode_kv_relayout_batch, P(E._ptr(self.cross_kv[l])),
                              P(E._ptr(self.cross_k_r[l])), P(E._ptr(self.cross_v_r[l])), B, Se)
        self.act = dict(xemb=zb(B, 2048), raw0=zb(B, d), raw1=zb(B, d), raw2=zb(B, d), hn=zb(B, d), qkv=zb(B, 3 * d),
                        qc=zb(B, d), ob=zb(B, d), f1=zb(B, F))
        self.barrier = torch.zeros(1, device=dev, dtype=torch.int32)
        self.err_flag = torch.zeros(1, device=dev, dtype=torch.int32)
        self.ln_stats = torch.zeros(3 * 64 * 2, device=dev, dtype=torch.float32)
        D = L.DecodeBatchDesc()
        self._fill_layers(D)
        D.n_layers, D.S_enc, D.S_max, D.B = nl, Se, S, B
        self._fill_common(D)
        D.enc_keep = E._ptr(eg.enc_keep)
        D.logits = E._ptr(self.logits)
        for k, t in self.act.items():
            setattr(D, k, E._ptr(t))
        D.barrier, D.error_flag, D.stats = E._ptr(self.barrier), E._ptr(self.err_flag), E._ptr(self.ln_stats)
        if os.environ.get('PIANOBART_B200_DECODE_SPLIT', '1') != '0':
            # scratch for the key halves of the cross-attention units of the last, partial round (csrc/decode_batch.cu)
            self.attn_part = torch.zeros(160 * 132, device=dev, dtype=torch.float32)
            self.attn_part_cnt = torch.zeros(80, device=dev, dtype=torch.int32)
            D.part, D.part_cnt = E._ptr(self.attn_part), E._ptr(self.attn_part_cnt)
        self.pdesc = D
        self.launches_per_step = 1
        self._steps_issued = 0

    def _fill_layers(self, D):
        for l in range(self.pb.layout.dec_layers):
            lp = 'bart.decoder.layers.%d' % l
            sa, ca = lp + '.self_attn', lp + '.encoder_attn'
            Ld = D.layer[l]
            Ld.wqkv, Ld.bqkv = self._W(sa + '.wqkv'), self._Pf(sa + '.bqkv')
            Ld.wo, Ld.bo = self._W(sa + '.out_proj.weight'), self._Pf(sa + '.out_proj.bias')
            Ld.ln1_g, Ld.ln1_b = self._Pf(lp + '.self_attn_layer_norm.weight'), self._Pf(lp + '.self_attn_layer_norm.bias')
            Ld.wqc, Ld.bqc = self._W(ca + '.q_proj.weight'), self._Pf(ca + '.q_proj.bias')
            Ld.woc, Ld.boc = self._W(ca + '.out_proj.weight'), self._Pf(ca + '.out_proj.bias')
            Ld.ln2_g, Ld.ln2_b = self._Pf(lp + '.encoder_attn_layer_norm.weight'), self._Pf(lp + '.encoder_attn_layer_norm.bias')
            Ld.w1, Ld.b1 = self._W(lp + '.fc1.weight'), self._Pf(lp + '.fc1.bias')
            Ld.w2, Ld.b2 = self._W(lp + '.fc2.weight'), self._Pf(lp + '.fc2.bias')
            Ld.ln3_g, Ld.ln3_b = self._Pf(lp + '.final_layer_norm.weight'), self._Pf(lp + '.final_layer_norm.bias')
            Ld.self_k, Ld.self_v = E._ptr(self.self_k_r[l]), E._ptr(self.self_v_r[l])
            Ld.cross_k, Ld.cross_v = E._ptr(self.cross_k_r[l]), E._ptr(self.cross_v_r[l])

    def _fill_common(self, D):
        D.emb_table = self._W('emb')
        D.w_in, D.b_in = self._W('encoder_linear.weight'), self._Pf('encoder_linear.bias')
        D.pos_table = self._W('bart.decoder.embed_positions.weight')
        D.lne_g, D.lne_b = self._Pf('bart.decoder.layernorm_embedding.weight'), self._Pf('bart.decoder.layernorm_embedding.bias')
        D.w_heads, D.b_heads = self._W('heads.w'), self._Pf('heads.b')
        D.t_dev, D.cur_tok, D.result, D.sampled = E._ptr(self.t_dev), E._ptr(self.cur_tok), E._ptr(self.result), E._ptr(self.sampled)
        D.done, D.n_written, D.uniforms, D.forced = E._ptr(self.done), E._ptr(self.n_written), E._ptr(self.uniforms), None

    def _build(self):
        if self.persist_batch:
            return self._build_persist_batch()
        if self.persist:
            return self._build_persist()
        if self.B <= 8 and not self.force_gemm:
            return self._build_small_batch()
        pb, lay = self.pb, self.pb.layout
        B, d, F, Se, S = self.B, self.d, self.F, self.Se, self.S
        eg = self.enc_graph
        # ---- prefill: cross-attention K/V of every decoder layer from the encoder output
        for l in range(lay.dec_layers):
            ca = 'bart.decoder.layers.%d.encoder_attn' % l
            self.prefill.gemm(E._ptr(eg.enc_out), self._W(ca + '.wkv'), E._ptr(self.cross_kv[l]), B * Se, 2 * d, d, d, d, 2 * d,
                              bias=self._Pf(ca + '.bkv'), name='kv_c%d' % l)
        # ---- one decode step
        st = self.step
        P = C.c_void_p
        st._add('embed', self.lib.pb_octuple_embed_fwd, P(E._ptr(self.cur_tok)), 0, P(self._W('emb')), P(E._ptr(self.x_emb)),
                C.c_longlong(B), self.ntok_arr, E.PB_BF16, P(None))
        self._skinny(st, self.x_emb, self._W('encoder_linear.weight'), d, 2048, 'in_linear')
        self._finalize(st, d, bias=self._Pf('encoder_linear.bias'), pos=self._W('bart.decoder.embed_positions.weight'),
                       ln='bart.decoder.layernorm_embedding', out=self.hA)
        h = self.hA
        for l in range(lay.dec_layers):
            lp = 'bart.decoder.layers.%d' % l
            sa, ca = lp + '.self_attn', lp + '.encoder_attn'
            self._skinny(st, h, self._W(sa + '.wqkv'), 3 * d, d, 'qkv%d' % l)
            self._finalize(st, 3 * d, bias=self._Pf(sa + '.bqkv'), out=self.qkv)
            cache = self.self_cache[l]
            self._attn(st, E._ptr(self.qkv), 3 * d, E._ptr(self.qkv, d), E._ptr(self.qkv, 2 * d), E._ptr(cache), E._ptr(cache, d),
                       S * 2 * d, 2 * d, 0, 0, 1, self.ao, S)
            self._skinny(st, self.ao, self._W(sa + '.out_proj.weight'), d, d, 'o%d' % l)
            self._finalize(st, d, bias=self._Pf(sa + '.out_proj.bias'), residual=h, ln=lp + '.self_attn_layer_norm', out=self.h1)
            self._skinny(st, self.h1, self._W(ca + '.q_proj.weight'), d, d, 'qc%d' % l)
            self._finalize(st, d, bias=self._Pf(ca + '.q_proj.bias'), out=self.qc)
            kv = self.cross_kv[l]
            self._attn(st, E._ptr(self.qc), d, 0, 0, E._ptr(kv), E._ptr(kv, d), Se * 2 * d, 2 * d, E._ptr(eg.enc_keep), Se, 0,
                       self.ao, Se)
            self._skinny(st, self.ao, self._W(ca + '.out_proj.weight'), d, d, 'oc%d' % l)
            self._finalize(st, d, bias=self._Pf(ca + '.out_proj.bias'), residual=self.h1, ln=lp + '.encoder_attn_layer_norm',
                           out=self.h2)
            self._skinny(st, self.h2, self._W(lp + '.fc1.weight'), F, d, 'fc1_%d' % l)
            self._finalize(st, F, bias=self._Pf(lp + '.fc1.bias'), out=self.f1, gelu=1)
            self._skinny(st, self.f1, self._W(lp + '.fc2.weight'), d, F, 'fc2_%d' % l)
            nxt = self.hB if h is self.hA else self.hA
            self._finalize(st, d, bias=self._Pf(lp + '.fc2.bias'), residual=self.h2, ln=lp + '.final_layer_norm', out=nxt)
            h = nxt
        self._skinny(st, h, self._W('heads.w'), E.VOCAB, d, 'heads')
        self._finalize(st, E.VOCAB, bias=self._Pf('heads.b'), out_f32=self.logits)
        self._sample_idx = len(st.ops)
        st._add('sample', self.lib.pb_decode_sample, P(E._ptr(self.logits)), P(E._ptr(self.uniforms)), P(None),
                P(E._ptr(self.t_dev)), P(E._ptr(self.cur_tok)), P(E._ptr(self.sampled)), B, S, self.ntok_arr, self.temp_arr,
                self.p_arr)
        st._add('advance', self.lib.pb_decode_advance, P(E._ptr(self.cur_tok)), P(E._ptr(self.result)), P(E._ptr(self.done)),
                P(E._ptr(self.t_dev)), P(E._ptr(self.n_written)), B, S, self.pad_arr)
        self.launches_per_step = len(st.ops)

    def _set_forced(self, forced):
        """Teacher forcing (parity tests): the token fed to step t+1 is forced[b, t] instead of the sampled one."""
        P = C.c_void_p
        if self.persist or self.persist_batch:
            self.forced = None if forced is None else forced.to(device=self.dev, dtype=torch.int32).contiguous()
            self.pdesc.forced = None if forced is None else E._ptr(self.forced)
            if self.persist:
                self.pdesc.stop_when_done = 1 if forced is None else 0
            return
        name, fn, args = self.step.ops[self._sample_idx]
        args = list(args)
        if forced is None:
            self.forced = None
            args[2] = P(None)
        else:
            self.forced = forced.to(device=self.dev, dtype=torch.int32).contiguous()
            args[2] = P(E._ptr(self.forced))
        self.step.ops[self._sample_idx] = (name, fn, tuple(args))
        self.graph = None

    # ------------------------------------------------------------------
    def start(self, input_ids_encoder, encoder_attention_mask, uniforms=None, forced=None):
        pb = self.pb
        pb.check_pack_generation(self._pack_gen, 'Generator')
        pb._sync_weights()
        pb._live_graph = None
        if (forced is None) != (self.forced is None) or forced is not None:
            self._set_forced(forced)
        eg = self.enc_graph
        eg.set_inputs(input_ids_encoder, encoder_attention_mask)
        eg.forward()
        self.prefill.run()
        self.t_dev.zero_()
        self.done.zero_()
        self.n_written.zero_()
        self.acc.zero_()
        self.cur_tok.copy_(self.sos.unsqueeze(0).expand(self.B, 8))
        pad = torch.tensor(pb.pad_word_np, dtype=torch.int32, device=self.dev)
        self.result.copy_(pad.view(1, 1, 8).expand_as(self.result))
        if uniforms is not None:
            self.uniforms.copy_(torch.as_tensor(uniforms, dtype=torch.float64).reshape(self.B, self.S, 8), non_blocking=True)
        if self.persist or self.persist_batch:
            self.err_flag.zero_()
            self._steps_issued = 0
        if self.persist_batch:
            self.ln_stats.zero_()

    def _run_persist(self, n):
        n = min(n, self.S - self._steps_issued)
        if n <= 0:
            return 0
        fn = self.lib.pb_decode_batch_run if self.persist_batch else self.lib.pb_decode_persist_run
        L.check(fn(C.byref(self.pdesc), n, self.ntok_arr, self.temp_arr, self.p_arr, self.pad_arr, L.stream_ptr()), 'decode_persist')
        self._steps_issued += n
        self.launches_per_step = 1.0 / n
        return 1

    def run_steps(self, n):
        """Runs n decode steps: one cooperative launch (batch 1, persistent kernel) or n CUDA-graph replays.
        Returns the number of library launches issued."""
        if self.persist or self.persist_batch:
            return self._run_persist(n)
        if not self.use_graph:
            for _ in range(n):
                self.step.run()
            return n * self.launches_per_step
        if self.graph is None:
            # one eager step to warm up (function attributes, tensor-map cache), then rewind its side effects
            snap = [t.clone() for t in (self.t_dev, self.cur_tok, self.result, self.done, self.n_written, self.sampled)]
            caches = [c[:, :1].clone() for c in self.self_cache]
            self.step.run()
            torch.cuda.synchronize()
            for t, s in zip((self.t_dev, self.cur_tok, self.result, self.done, self.n_written, self.sampled), snap):
                t.copy_(s)
            for c, s in zip(self.self_cache, caches):
                c[:, :1].copy_(s)
            self.acc.zero_()
            # the per-token kernels are a few microseconds each: programmatic dependent launch edges inside the graph
            # measured slower than plain kernel-to-kernel edges (profiles/r1_summary.md), so the capture turns them off
            lib = L.lib()
            prev = lib.pb_set_pdl(1 if os.environ.get('PIANOBART_B200_DECODE_PDL', '0') == '1' else 0)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.step.run()
            finally:
                lib.pb_set_pdl(prev)
            self.graph = g
            # capture does not execute: state is unchanged
        for _ in range(n):
            self.graph.replay()
        return n * self.launches_per_step

    def finish(self):
        torch.cuda.synchronize()
        if (self.persist or self.persist_batch) and int(self.err_flag.item()) != 0:
            raise L.PBError('decode_persist_kernel reported error %d' % int(self.err_flag.item()))
        return self.result.to(torch.int64), self.n_written.cpu().numpy(), self.done.cpu().numpy()


_GEN_CACHE = {}


def generate(lm, input_ids_encoder, encoder_attention_mask=None, check_every=32):
    """Drop-in for `PianoBartLM.forward(..., generate=True)`; consumes numpy's global RNG like the reference
    (8 uniforms per executed step, attribute order)."""
    B, S = input_ids_encoder.shape[0], input_ids_encoder.shape[1]
    lm.pianobart._ensure_packed()
    key = (id(lm), B, S, lm.pianobart._pack_gen)
    gen = _GEN_CACHE.get(key)
    if gen is None:
        _GEN_CACHE.clear()
        gen = Generator(lm, B, S, S)
        _GEN_CACHE[key] = gen
    state = np.random.get_state()
    uniforms = np.random.random_sample((B, S, 8))
    with torch.no_grad():
        gen.start(input_ids_encoder, encoder_attention_mask, uniforms)
        done_steps = 0
        while done_steps < S:
            n = min(check_every, S - done_steps)
            gen.run_steps(n)
            done_steps += n
            if bool(gen.done.all().item()):
                break
        result, n_written, done = gen.finish()
    if B == 1:
        # leave numpy's global stream exactly where the reference would: 8 draws per executed step
        executed = int(n_written[0]) + (1 if done[0] else 0)
        np.random.set_state(state)
        if executed:
            np.random.random_sample(8 * executed)
    return result.to(input_ids_encoder.device)

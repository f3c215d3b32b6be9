"""Finetuning classifier heads on the kernel path (kernels K15 / K16).

Autograd bridges over the C ABI for everything the reference evaluates around the backbone in
`SequenceClassification` (model.py:128-143,195-218) and `TokenClassification` (model.py:236-272,
replacement decoder front end PianoBart.py:9-16 + finetune.py:194-198) and for the trainer's loss
(finetune.py:125-132,233-235):

    linear()          nn.Linear: wide outputs -> pb_gemm_bf16 / pb_gemm_f32 (+ wgrad / dgrad / pb_colsum),
                      <= 16 outputs -> pb_smalln_linear_* with the preceding ReLU / tanh folded into the operand
    dropout()         nn.Dropout -> pb_dropout_apply (counter-based mask, regenerated in backward)
    seq_softmax()     softmax over the sequence axis -> pb_seq_softmax_*
    attn_pool()       torch.bmm(attn_mat, x) -> pb_attn_pool_*
    embed_rows()      Embeddings.forward -> pb_rows_gather / pb_rows_scatter_add
    masked_ce()       CrossEntropyLoss(reduction='none') * mask / sum(mask) (+ argmax, #correct) -> pb_heads_ce

Activations between these ops are fp32 tensors (the heads are < 0.1 % of the step); in bf16 mode GEMM operands are cast with
pb_cast_from_f32 and accumulate / output in fp32.  There is no PyTorch fallback: the functions raise on CPU tensors.
"""
import ctypes as C

import torch

from . import _lib as L
from . import engine as E

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
SMALL_N = 16


def _p(t):
    return C.c_void_p(t.data_ptr())


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t):
    if not t.is_cuda:
        raise L.PBError('pianobart_b200 heads have no CPU path (sm_100a kernels only)')


def _f32(t):
    return t.detach().contiguous().float()


def _cast(t32, pb_dtype):
    """fp32 tensor -> GEMM operand dtype (own cast kernel; identity in fp32 mode)"""
    if pb_dtype == E.PB_F32:
        return t32
    out = torch.empty(t32.shape, dtype=torch.bfloat16, device=t32.device)
    L.check(L.lib().pb_cast_from_f32(_p(t32), _p(out), C.c_longlong(t32.numel()), C.c_float(1.0), pb_dtype, _s()), 'cast')
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, pb_dtype, act_in):
        _need_cuda(x)
        x2, w2 = _f32(x).view(-1, x.shape[-1]), _f32(w)
        M, K, N = x2.shape[0], x2.shape[1], w2.shape[0]
        b2 = None if b is None else _f32(b)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        lib = L.lib()
        small = N <= SMALL_N
        if act_in != ACT_NONE and not small:
            raise L.PBError('heads.linear: a folded input activation needs <= %d outputs' % SMALL_N)
        if small:
            L.check(lib.pb_smalln_linear_fwd(_p(x2), _p(w2), _p(b2) if b2 is not None else None, _p(y), C.c_longlong(M), N, K,
                                             act_in, _s()), 'smalln_linear_fwd')
            ctx.save_for_backward(x2, w2)
        else:
            xo, wo = _cast(x2, pb_dtype), _cast(w2, pb_dtype)
            plan = E.Plan(pb_dtype)
            plan.gemm(xo.data_ptr(), wo.data_ptr(), y.data_ptr(), M, N, K, K, K, N, bias=0 if b2 is None else b2.data_ptr(),
                      flags=L.PB_GEMM_OUT_F32, name='head.linear')
            plan.run()
            ctx.save_for_backward(xo, wo)
        ctx.cfg = (pb_dtype, act_in, small, M, N, K, b is not None, x.shape)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        pb_dtype, act_in, small, M, N, K, has_b, xshape = ctx.cfg
        xs, ws = ctx.saved_tensors
        dy2 = _f32(dy).view(M, N)
        lib = L.lib()
        dev = dy.device
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty(M, K, dtype=torch.float32, device=dev) if need_dx else None
        dw = torch.zeros(N, K, dtype=torch.float32, device=dev)
        db = torch.zeros(N, dtype=torch.float32, device=dev) if has_b else None
        if small:
            L.check(lib.pb_smalln_linear_bwd(_p(xs), _p(ws), _p(dy2), _p(dx) if need_dx else None, _p(dw),
                                             _p(db) if has_b else None, C.c_longlong(M), N, K, act_in, _s()),
                    'smalln_linear_bwd')
        else:
            dyo = _cast(dy2, pb_dtype)
            plan = E.Plan(pb_dtype)
            if need_dx:
                plan.gemm(dyo.data_ptr(), ws.data_ptr(), dx.data_ptr(), M, K, N, N, K, K, b_mn=1, flags=L.PB_GEMM_OUT_F32,
                          name='head.dgrad')
            plan.wgrad(dyo.data_ptr(), xs.data_ptr(), dw.data_ptr(), N, K, M, N, K, name='head.wgrad')
            if has_b:
                plan.colsum(dyo.data_ptr(), db.data_ptr(), M, N, N)
            plan.run()
        return (dx.view(xshape) if need_dx else None), dw, db, None, None


def linear(x, weight, bias, pb_dtype, act_in=ACT_NONE):
    """y = act_in(x) W^T + b over the last axis of x."""
    return _LinearFn.apply(x, weight, bias, pb_dtype, act_in)


class DropSeeds:
    """Device-resident seeds for the heads' dropout sites: slot k of a pre-filled table serves the k-th training forward,
    so neither a host->device copy nor a bump kernel sits between the steps, and backward reads the slot its forward used."""

    SLOTS = 1024

    def __init__(self, device):
        base = torch.initial_seed() & 0x3fffffffffffffff
        self.table = (torch.arange(self.SLOTS, dtype=torch.int64) + base).to(device)
        self.count = 0

    def next_site(self, p):
        k = self.count
        self.count += 1
        st = L.DropSite()
        st.seed = self.table.data_ptr() + 8 * (k % self.SLOTS)
        st.op = 0x4000 + (k // self.SLOTS) % 0x4000
        st.thresh = min(int((1.0 - p) * 4294967296.0), 4294967295)
        st.scale = 1.0 / (1.0 - p)
        return st


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, site, seeds):
        _need_cuda(x)
        x2 = _f32(x)
        y = torch.empty_like(x2)
        L.check(L.lib().pb_dropout_apply(_p(x2), _p(y), C.c_longlong(x2.numel()), C.byref(site), _s()), 'dropout_apply')
        ctx.site, ctx.seeds = site, seeds      # (seeds keeps the table alive)
        return y

    @staticmethod
    def backward(ctx, dy):
        d2 = _f32(dy)
        dx = torch.empty_like(d2)
        L.check(L.lib().pb_dropout_apply(_p(d2), _p(dx), C.c_longlong(d2.numel()), C.byref(ctx.site), _s()), 'dropout_apply')
        return dx, None, None


def dropout(x, p, training, seeds):
    """nn.Dropout(p): identity in eval mode."""
    if not training or p <= 0.0:
        return x
    return _DropoutFn.apply(x, seeds.next_site(p), seeds)


class _SeqSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        _need_cuda(a)
        a3 = _f32(a)
        B, S, R = a3.shape
        p = torch.empty_like(a3)
        L.check(L.lib().pb_seq_softmax_fwd(_p(a3), _p(p), B, S, R, _s()), 'seq_softmax_fwd')
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, dp):
        p, = ctx.saved_tensors
        B, S, R = p.shape
        d2 = _f32(dp)
        da = torch.empty_like(p)
        L.check(L.lib().pb_seq_softmax_bwd(_p(p), _p(d2), _p(da), B, S, R, _s()), 'seq_softmax_bwd')
        return da


def seq_softmax(a):
    """softmax(a, dim=1) for a [B, S, R] tensor, R <= 8 (model.py:142)."""
    return _SeqSoftmaxFn.apply(a)


class _AttnPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, x):
        _need_cuda(x)
        p3, x3 = _f32(p), _f32(x)
        B, S, R = p3.shape
        D = x3.shape[2]
        m = torch.empty(B, R, D, dtype=torch.float32, device=x.device)
        L.check(L.lib().pb_attn_pool_fwd(_p(p3), _p(x3), _p(m), B, S, R, D, _s()), 'attn_pool_fwd')
        ctx.save_for_backward(p3, x3)
        return m

    @staticmethod
    def backward(ctx, dm):
        p3, x3 = ctx.saved_tensors
        B, S, R = p3.shape
        D = x3.shape[2]
        d2 = _f32(dm)
        dx, dp = torch.empty_like(x3), torch.empty_like(p3)
        L.check(L.lib().pb_attn_pool_bwd(_p(p3), _p(x3), _p(d2), _p(dx), _p(dp), B, S, R, D, _s()), 'attn_pool_bwd')
        return dp, dx


def attn_pool(p, x):
    """m[b, r, :] = sum_s p[b, s, r] x[b, s, :]  == torch.bmm(p.permute(0, 2, 1), x)  (model.py:143,209)."""
    return _AttnPoolFn.apply(p, x)


class _EmbedRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table, scale):
        _need_cuda(table)
        ids2 = ids.detach().contiguous().long()
        t2 = _f32(table)
        n_rows, d = t2.shape
        out = torch.empty(*ids2.shape, d, dtype=torch.float32, device=table.device)
        err = torch.zeros(1, dtype=torch.int32, device=table.device)
        L.check(L.lib().pb_rows_gather(_p(ids2), _p(t2), _p(out), C.c_longlong(ids2.numel()), n_rows, d, C.c_float(scale),
                                       _p(err), _s()), 'rows_gather')
        ctx.save_for_backward(ids2)
        ctx.cfg = (n_rows, d, scale, err)
        return out

    @staticmethod
    def backward(ctx, dout):
        ids2, = ctx.saved_tensors
        n_rows, d, scale, _ = ctx.cfg
        d2 = _f32(dout)
        dt = torch.zeros(n_rows, d, dtype=torch.float32, device=dout.device)
        L.check(L.lib().pb_rows_scatter_add(_p(ids2), _p(d2), _p(dt), C.c_longlong(ids2.numel()), n_rows, d, C.c_float(scale),
                                            _s()), 'rows_scatter_add')
        return None, dt, None


def embed_rows(ids, table, scale):
    """table[ids] * scale (Embeddings.forward, PianoBart.py:15-16)."""
    return _EmbedRowsFn.apply(ids, table, scale)


class _MaskedCEFn(torch.autograd.Function):
    """loss = sum_m CE(logits[m], target[m]) mask[m] / den with the trainer's argmax / #correct from the same pass."""

    @staticmethod
    def forward(ctx, logits, target, mask, den):
        _need_cuda(logits)
        lg = _f32(logits)
        M, Cn = lg.shape
        dev = lg.device
        tg = target.detach().reshape(M, 1).to(torch.int32).contiguous()
        mk = mask.detach().reshape(M, 1).float().contiguous()
        dn = den.detach().reshape(1).float().contiguous()
        loss_num = torch.zeros(1, dtype=torch.float32, device=dev)
        correct = torch.zeros(1, dtype=torch.float32, device=dev)
        dlogits = torch.empty_like(lg)
        argmax = torch.empty(M, 1, dtype=torch.int32, device=dev)
        seg = (C.c_int * 1)(Cn)
        w = (C.c_float * 1)(1.0)
        L.check(L.lib().pb_heads_ce(_p(lg), _p(tg), _p(mk), _p(dn), _p(loss_num), _p(correct), _p(dlogits), _p(argmax),
                                    C.c_longlong(M), 1, seg, w, C.c_float(1.0), E.PB_F32, _s()), 'heads_ce')
        ctx.save_for_backward(dlogits)
        ctx.mark_non_differentiable(correct, argmax)
        return (loss_num / dn).squeeze(0), correct.squeeze(0), argmax.view(M)

    @staticmethod
    def backward(ctx, g_loss, g_correct, g_argmax):
        dlogits, = ctx.saved_tensors
        return dlogits * g_loss, None, None, None


def masked_ce(logits, target, mask=None):
    """logits [M, C] (C <= 16 ... any), target [M], mask [M] or None -> (loss, #correct (masked), argmax [M]).
    mask None: plain mean over the M rows (finetune.py:131-132)."""
    M = logits.shape[0]
    if mask is None:
        mask = torch.ones(M, dtype=torch.float32, device=logits.device)
        den = torch.full((1,), float(M), dtype=torch.float32, device=logits.device)
    else:
        mask = mask.reshape(M).float().contiguous()
        den = torch.zeros(1, dtype=torch.float32, device=logits.device)
        L.check(L.lib().pb_mask_sums(_p(mask), _p(den), C.c_longlong(M), 1, _s()), 'mask_sums')
    return _MaskedCEFn.apply(logits, target, mask, den)

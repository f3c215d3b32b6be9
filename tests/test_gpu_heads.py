"""Op-level parity of the finetuning classifier-head kernels (csrc/cls_heads.cu through pianobart_b200/heads.py) against the
PyTorch fp32 ops the reference evaluates (model.py:128-143,173-178,195-218,244-272; finetune.py:125-132), forward and
backward, at the finetune configs' sizes (B 8, S 1024, hs 1024, da 128, r 4, 256 hidden, class_num 4 / 8)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _leaf(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda().requires_grad_(True)


@pytest.mark.parametrize('act', [0, 1, 2])
@pytest.mark.parametrize('shape', [(8192, 8, 256), (8192, 4, 128), (8, 4, 256), (77, 16, 100)])
def test_small_projection_fwd_bwd(act, shape):
    from pianobart_b200 import heads, engine as E
    M, N, K = shape
    x, w, b = _leaf(M, K, seed=1), _leaf(N, K, scale=0.1, seed=2), _leaf(N, seed=3)
    y = heads.linear(x, w, b, E.PB_F32, act_in=act)
    g = torch.randn(M, N, generator=torch.Generator().manual_seed(4)).cuda()
    y.backward(g)
    xr, wr, br = [t.detach().clone().requires_grad_(True) for t in (x, w, b)]
    a = {0: xr, 1: torch.relu(xr), 2: torch.tanh(xr)}[act]
    yr = F.linear(a, wr, br)
    yr.backward(g)
    assert _rel(y, yr) < 1e-5
    assert _rel(x.grad, xr.grad) < 1e-5
    assert _rel(w.grad, wr.grad) < 2e-5
    assert _rel(b.grad, br.grad) < 2e-5


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
@pytest.mark.parametrize('shape', [(8192, 256, 1024, True), (8192, 128, 1024, False), (8, 256, 4096, True),
                                   (8192, 1024, 64, True), (5, 256, 4096, True)])
def test_wide_linear_fwd_bwd(dtype, shape):
    """nn.Linear through the tcgen05 (bf16) / SIMT (fp32) GEMMs incl. the M = batch-size rows of the pooled classifier"""
    from pianobart_b200 import heads, engine as E
    M, N, K, bias = shape
    pbd = E.PB_F32 if dtype == 'fp32' else E.PB_BF16
    x, w = _leaf(M, K, seed=5), _leaf(N, K, scale=K ** -0.5, seed=6)
    b = _leaf(N, seed=7) if bias else None
    y = heads.linear(x, w, b, pbd)
    g = torch.randn(M, N, generator=torch.Generator().manual_seed(8)).cuda()
    y.backward(g)
    xr, wr = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    br = b.detach().clone().requires_grad_(True) if bias else None
    yr = F.linear(xr, wr, br)
    yr.backward(g)
    tol = 2e-5 if dtype == 'fp32' else 2e-2
    assert _rel(y, yr) < tol
    assert _rel(x.grad, xr.grad) < tol
    assert _rel(w.grad, wr.grad) < tol
    if bias:
        assert _rel(b.grad, br.grad) < tol


def test_seq_softmax_and_pool_fwd_bwd():
    from pianobart_b200 import heads
    B, S, R, D = 8, 1024, 4, 1024
    a, x = _leaf(B, S, R, scale=2.0, seed=9), _leaf(B, S, D, seed=10)
    p = heads.seq_softmax(a)
    m = heads.attn_pool(p, x)
    g = torch.randn(B, R, D, generator=torch.Generator().manual_seed(11)).cuda()
    m.backward(g)
    ar, xr = a.detach().clone().requires_grad_(True), x.detach().clone().requires_grad_(True)
    pr = torch.softmax(ar, dim=1)
    mr = torch.bmm(pr.permute(0, 2, 1), xr)
    mr.backward(g)
    assert _rel(p, pr) < 1e-5
    assert _rel(m, mr) < 1e-5
    assert _rel(x.grad, xr.grad) < 1e-5
    assert _rel(a.grad, ar.grad) < 1e-4
    # ragged shapes: S not a multiple of the pooling chunk, D not a multiple of the block, R = 8
    a, x = _leaf(3, 333, 8, seed=12), _leaf(3, 333, 200, seed=13)
    m = heads.attn_pool(heads.seq_softmax(a), x)
    m.sum().backward()
    ar, xr = a.detach().clone().requires_grad_(True), x.detach().clone().requires_grad_(True)
    mr = torch.bmm(torch.softmax(ar, dim=1).permute(0, 2, 1), xr)
    mr.sum().backward()
    assert _rel(m, mr) < 1e-5 and _rel(x.grad, xr.grad) < 1e-5
    assert float((a.grad - ar.grad).abs().max()) < 1e-5


def test_dropout_matches_its_mask_and_backward_reuses_it():
    from pianobart_b200 import heads, _lib as L
    x = _leaf(8, 4096, seed=14)
    seeds = heads.DropSeeds(x.device)
    y = heads.dropout(x, 0.1, True, seeds)
    y.backward(torch.ones_like(y))
    site = L.DropSite()
    site.seed, site.op = seeds.table.data_ptr(), 0x4000
    site.thresh = min(int(0.9 * 4294967296.0), 4294967295)
    mask = torch.empty(x.numel(), dtype=torch.uint8, device='cuda')
    L.check(L.lib().pb_dropout_mask(C.c_void_p(site.seed), site.op, site.thresh, C.c_void_p(mask.data_ptr()),
                                    C.c_longlong(x.numel()), L.stream_ptr()), 'dropout_mask')
    keep = mask.view_as(x).float()
    assert torch.equal(y.detach(), x.detach() * keep * (1.0 / 0.9))
    assert torch.equal(x.grad, keep * (1.0 / 0.9))
    assert 0.88 < float(keep.mean()) < 0.92
    # second call: another slot, another mask; eval mode: identity
    y2 = heads.dropout(x, 0.1, True, seeds)
    assert not torch.equal(y2 == 0, y == 0)
    assert heads.dropout(x, 0.1, False, seeds) is x


def test_embed_rows_fwd_bwd():
    from pianobart_b200 import heads
    table = _leaf(8, 64, seed=15)
    ids = torch.randint(0, 8, (8, 1024), generator=torch.Generator().manual_seed(16)).cuda()
    out = heads.embed_rows(ids, table, 8.0)
    g = torch.randn(8, 1024, 64, generator=torch.Generator().manual_seed(17)).cuda()
    out.backward(g)
    tr = table.detach().clone().requires_grad_(True)
    outr = F.embedding(ids, tr) * 8.0
    outr.backward(g)
    assert torch.equal(out.detach(), outr.detach())
    assert _rel(table.grad, tr.grad) < 1e-5


@pytest.mark.parametrize('Cn', [4, 8])
def test_masked_ce_matches_torch(Cn):
    from pianobart_b200 import heads
    M = 8 * 1024
    lg = _leaf(M, Cn, scale=3.0, seed=18)
    tg = torch.randint(0, Cn, (M,), generator=torch.Generator().manual_seed(19)).cuda()
    mk = (torch.rand(M, generator=torch.Generator().manual_seed(20)) < 0.8).float().cuda()
    loss, correct, am = heads.masked_ce(lg, tg, mk)
    (loss * 0.5).backward()
    lr = lg.detach().clone().requires_grad_(True)
    ce = F.cross_entropy(lr, tg, reduction='none')
    lossr = (ce * mk).sum() / mk.sum()
    (lossr * 0.5).backward()
    assert abs(float(loss.detach()) - float(lossr.detach())) < 1e-5 * abs(float(lossr.detach()))
    assert _rel(lg.grad, lr.grad) < 1e-5
    assert torch.equal(am.long(), lr.argmax(-1))
    assert float(correct) == float(((lr.argmax(-1) == tg).float() * mk).sum())
    # sequence tasks: plain mean
    loss2, correct2, _ = heads.masked_ce(lg[:8], tg[:8], None)
    assert abs(float(loss2) - float(F.cross_entropy(lr[:8], tg[:8]))) < 1e-5


def test_head_optimizer_matches_hf_adamw_semantics():
    """HFAdamW (one pb_adamw launch per head tensor) against the transformers-4.29 update rule written out in fp64"""
    from pianobart_b200.finetune import HFAdamW
    p = torch.nn.Parameter(torch.randn(256, 40, generator=torch.Generator().manual_seed(21)).cuda())
    opt = HFAdamW([p], lr=1e-3, weight_decay=0.01)
    ref = p.detach().double().cpu().numpy().copy()
    m = np.zeros_like(ref); v = np.zeros_like(ref)
    for step in range(1, 4):
        g = torch.randn(256, 40, generator=torch.Generator().manual_seed(30 + step)).cuda()
        p.grad = g.clone()
        opt.step()
        gn = g.double().cpu().numpy()
        m = 0.9 * m + 0.1 * gn
        v = 0.999 * v + 0.001 * gn * gn
        ss = 1e-3 * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
        ref = ref - ss * m / (np.sqrt(v) + 1e-6)
        ref = ref - 1e-3 * 0.01 * ref
    assert np.abs(p.detach().double().cpu().numpy() - ref).max() < 1e-6


# ------------------------------------------------------------------ north-star fusion 3: MLM heads + masked CE in one kernel
@pytest.mark.parametrize('M,K', [(16 * 1024, 1024), (300, 64), (128, 128), (1, 64)])
def test_fused_heads_ce_matches_gemm_plus_ce_and_torch(M, K):
    """pb_heads_ce_fused (logits only ever in TMEM) against (a) the unfused library path pb_gemm_bf16 + pb_heads_ce and
    (b) torch fp32 cross-entropy on the same bf16 operands (model.py:119-126 + pretrain.py:112-118,163-189)."""
    from pianobart_b200 import _lib as L, engine as E
    lib = L.lib()
    P = C.c_void_p
    V, seg_sizes = E.VOCAB, E.N_TOKENS
    g = torch.Generator().manual_seed(100 + M)
    h = (torch.randn(M, K, generator=g) * 1.0).to(torch.bfloat16).cuda()
    w = (torch.randn(V, K, generator=g) * (2.0 / K ** 0.5)).to(torch.bfloat16).cuda()
    bias = (torch.randn(V, generator=g) * 0.5).cuda()
    tg = torch.stack([torch.randint(0, n, (M,), generator=g) for n in seg_sizes], dim=1).to(torch.int32).cuda()
    mk = (torch.rand(M, 8, generator=g) < 0.3).float().cuda()
    if M > 4:
        mk[3] = 0.0
    den = torch.zeros(8, device='cuda')
    L.check(lib.pb_mask_sums(P(mk.data_ptr()), P(den.data_ptr()), C.c_longlong(M), 8, L.stream_ptr()), 'mask_sums')
    den.clamp_(min=1.0)
    seg = (C.c_int * 8)(*seg_sizes)
    wts = (C.c_float * 8)(262, 134, 262, 134, 38, 135, 55, 260)

    loss_f, cor_f = torch.zeros(8, device='cuda'), torch.zeros(8, device='cuda')
    dl_f = torch.full((M, V), 7.0, dtype=torch.bfloat16, device='cuda')
    am_f = torch.full((M, 8), -1, dtype=torch.int32, device='cuda')
    L.check(lib.pb_heads_ce_fused(P(h.data_ptr()), C.c_longlong(K), P(w.data_ptr()), P(bias.data_ptr()), P(tg.data_ptr()),
                                  P(mk.data_ptr()), P(den.data_ptr()), P(loss_f.data_ptr()), P(cor_f.data_ptr()),
                                  P(dl_f.data_ptr()), P(am_f.data_ptr()), C.c_longlong(M), K, 8, seg, wts, C.c_float(0.7),
                                  L.stream_ptr()), 'heads_ce_fused')
    # (a) unfused library path
    logits = torch.empty(M, V, dtype=torch.float32, device='cuda')
    plan = E.Plan(E.PB_BF16)
    plan.gemm(h.data_ptr(), w.data_ptr(), logits.data_ptr(), M, V, K, K, K, V, bias=bias.data_ptr(), flags=L.PB_GEMM_OUT_F32)
    plan.run()
    loss_u, cor_u = torch.zeros(8, device='cuda'), torch.zeros(8, device='cuda')
    dl_u = torch.empty(M, V, dtype=torch.bfloat16, device='cuda')
    am_u = torch.empty(M, 8, dtype=torch.int32, device='cuda')
    L.check(lib.pb_heads_ce(P(logits.data_ptr()), P(tg.data_ptr()), P(mk.data_ptr()), P(den.data_ptr()), P(loss_u.data_ptr()),
                            P(cor_u.data_ptr()), P(dl_u.data_ptr()), P(am_u.data_ptr()), C.c_longlong(M), 8, seg, wts,
                            C.c_float(0.7), E.PB_BF16, L.stream_ptr()), 'heads_ce')
    torch.cuda.synchronize()
    assert _rel(loss_f, loss_u) < 2e-5
    assert float((am_f != am_u).float().mean()) < 1e-4          # (equal logits up to accumulation order: ties only)
    assert float((cor_f - cor_u).abs().max()) <= max(2.0, 1e-4 * M)
    d = (dl_f.float() - dl_u.float()).abs()
    scale = dl_u.float().abs().max()
    assert float(d.max()) <= 0.01 * float(scale) + 1e-12         # one bf16 rounding step of the largest entries
    # (b) torch fp32 on the same operands
    lg = (h.float() @ w.float().t() + bias).requires_grad_(True)
    total = 0.0
    off = 0
    for s, n in enumerate(seg_sizes):
        ce = F.cross_entropy(lg[:, off:off + n], tg[:, s].long(), reduction='none')
        num = (ce * mk[:, s]).sum()
        assert abs(float(loss_f[s]) - float(num)) <= 2e-3 * max(1.0, abs(float(num)))
        total = total + num / den[s] * wts[s] / 1280.0 * 0.7
        off += n
    total.backward()
    assert float((dl_f.float() - lg.grad).abs().max()) <= 0.01 * float(lg.grad.abs().max()) + 1e-12
    # evaluation mode: no dlogits, same statistics
    loss_e, cor_e = torch.zeros(8, device='cuda'), torch.zeros(8, device='cuda')
    L.check(lib.pb_heads_ce_fused(P(h.data_ptr()), C.c_longlong(K), P(w.data_ptr()), P(bias.data_ptr()), P(tg.data_ptr()),
                                  P(mk.data_ptr()), P(den.data_ptr()), P(loss_e.data_ptr()), P(cor_e.data_ptr()), P(None),
                                  P(None), C.c_longlong(M), K, 8, seg, wts, C.c_float(1.0), L.stream_ptr()), 'heads_ce_fused')
    assert _rel(loss_e, loss_f) < 1e-5 and torch.equal(cor_e, cor_f)

"""GPU bring-up check of the tcgen05 attention kernels against torch fp32 (torch is only the checker)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pianobart_b200 import _lib as L
lib = L.lib(); dev = 'cuda:0'; torch.manual_seed(0)
fails = 0

def run(B, H, Sq, Sk, causal, pad, fused_qkv, tag):
    global fails
    hd = 128; d = H * hd
    if fused_qkv:
        qkv = (torch.randn(B, Sq, 3 * d, device=dev) * 0.7).bfloat16()
        q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
        ldq = ldk = ldv = 3 * d
        dqkv = torch.zeros_like(qkv); dq, dk, dv = dqkv[..., :d], dqkv[..., d:2 * d], dqkv[..., 2 * d:]
    else:
        q = (torch.randn(B, Sq, d, device=dev) * 0.7).bfloat16()
        kv = (torch.randn(B, Sk, 2 * d, device=dev) * 0.7).bfloat16()
        k, v = kv[..., :d], kv[..., d:]
        ldq, ldk, ldv = d, 2 * d, 2 * d
        dq = torch.zeros_like(q); dkv = torch.zeros_like(kv); dk, dv = dkv[..., :d], dkv[..., d:]
    keep = torch.ones(B, Sk, device=dev, dtype=torch.uint8)
    if pad:
        keep = (torch.rand(B, Sk, device=dev) > 0.3).to(torch.uint8); keep[:, 0] = 1
    o = torch.zeros(B, Sq, d, device=dev, dtype=torch.bfloat16)
    do = (torch.randn(B, Sq, d, device=dev) * 0.5).bfloat16()
    lse = torch.zeros(B, H, Sq, device=dev); dvec = torch.zeros(B, H, Sq, device=dev)
    a = L.AttnDesc()
    a.q, a.k, a.v, a.o, a.dout = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr()
    a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.ldq, a.ldk, a.ldv, a.ldo, a.lddo = ldq, ldk, ldv, d, d
    a.lddq, a.lddk, a.lddv = ldq, ldk, ldv
    a.lse, a.dvec, a.key_keep = lse.data_ptr(), dvec.data_ptr(), keep.data_ptr()
    a.B, a.H, a.Sq, a.Sk, a.hd, a.causal, a.scale = B, H, Sq, Sk, hd, causal, hd ** -0.5
    rc = lib.pb_attn_fwd(C.byref(a), L.stream_ptr())
    if rc: print('FAIL fwd launch', lib.pb_last_error().decode()); fails += 1; return
    torch.cuda.synchronize()
    qf = q.float().view(B, Sq, H, hd).transpose(1, 2).detach().requires_grad_(True)
    kf = k.float().view(B, Sk, H, hd).transpose(1, 2).detach().requires_grad_(True)
    vf = v.float().view(B, Sk, H, hd).transpose(1, 2).detach().requires_grad_(True)
    s = (qf @ kf.transpose(-1, -2)) * hd ** -0.5
    allow = (keep != 0)[:, None, None, :].expand(B, H, Sq, Sk)
    if causal: allow = allow & torch.ones(Sq, Sk, dtype=torch.bool, device=dev).tril()
    pr = torch.softmax(s.masked_fill(~allow, float('-inf')), -1)
    ref = (pr @ vf).transpose(1, 2).reshape(B, Sq, d)
    e_o = ((o.float() - ref).abs().max() / ref.abs().max()).item()
    rc = lib.pb_attn_bwd(C.byref(a), L.stream_ptr())
    if rc: print('FAIL bwd launch', lib.pb_last_error().decode()); fails += 1; return
    torch.cuda.synchronize()
    ref.backward(do.float())
    gq = qf.grad.transpose(1, 2).reshape(B, Sq, d); gk = kf.grad.transpose(1, 2).reshape(B, Sk, d); gv = vf.grad.transpose(1, 2).reshape(B, Sk, d)
    e_q = ((dq.float() - gq).abs().max() / gq.abs().max()).item()
    e_k = ((dk.float() - gk).abs().max() / gk.abs().max()).item()
    e_v = ((dv.float() - gv).abs().max() / gv.abs().max()).item()
    ok = max(e_o, e_q, e_k, e_v) < 3e-2
    print('%s %-30s B=%d H=%d Sq=%d Sk=%d causal=%d pad=%d  o %.2e dq %.2e dk %.2e dv %.2e' % ('ok  ' if ok else 'FAIL', tag, B, H, Sq, Sk, causal, pad, e_o, e_q, e_k, e_v))
    if not ok: fails += 1

run(1, 1, 128, 128, 0, 0, True, 'single tile')
run(1, 2, 256, 256, 0, 0, True, 'two blocks')
run(2, 2, 384, 384, 0, 1, True, 'enc self + padding')
run(2, 2, 384, 384, 1, 1, True, 'dec self causal + padding')
run(2, 2, 256, 384, 0, 1, False, 'cross')
run(1, 2, 200, 200, 1, 1, True, 'tails (S % 128 != 0)')
run(1, 1, 96, 320, 0, 1, False, 'cross tails')

def bench(B, H, S, causal):
    hd = 128; d = H * hd
    qkv = (torch.randn(B, S, 3 * d, device=dev) * 0.7).bfloat16(); dqkv = torch.zeros_like(qkv)
    o = torch.zeros(B, S, d, device=dev, dtype=torch.bfloat16); do = torch.randn(B, S, d, device=dev).bfloat16()
    lse = torch.zeros(B, H, S, device=dev); dvec = torch.zeros(B, H, S, device=dev)
    keep = torch.ones(B, S, device=dev, dtype=torch.uint8)
    a = L.AttnDesc()
    a.q, a.k, a.v = qkv.data_ptr(), qkv.data_ptr() + d * 2, qkv.data_ptr() + 4 * d
    a.o, a.dout = o.data_ptr(), do.data_ptr()
    a.dq, a.dk, a.dv = dqkv.data_ptr(), dqkv.data_ptr() + 2 * d, dqkv.data_ptr() + 4 * d
    a.ldq = a.ldk = a.ldv = a.lddq = a.lddk = a.lddv = 3 * d; a.ldo = a.lddo = d
    a.lse, a.dvec, a.key_keep = lse.data_ptr(), dvec.data_ptr(), keep.data_ptr()
    a.B, a.H, a.Sq, a.Sk, a.hd, a.causal, a.scale = B, H, S, S, hd, causal, hd ** -0.5
    for fn, nm, mult in ((lib.pb_attn_fwd, 'fwd', 4), (lib.pb_attn_bwd, 'bwd', 10)):
        for _ in range(3): fn(C.byref(a), L.stream_ptr())
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn(C.byref(a), L.stream_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = mult * B * H * S * S * hd * (0.5 if causal else 1.0)
        print('bench %s B=%d H=%d S=%d causal=%d: %.3f ms  %.1f TFLOP/s (algorithmic)' % (nm, B, H, S, causal, ms, fl / ms / 1e9))
bench(16, 8, 1024, 0)
bench(16, 8, 1024, 1)
print('FAILS', fails)

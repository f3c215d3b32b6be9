"""Times the fused heads + CE kernel against the unfused pair (heads GEMM + heads_ce) at the pretraining shape."""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
from pianobart_b200 import _lib as L, engine as E

lib = L.lib()
P = C.c_void_p
M, K, V = 16 * 1024, 1024, E.VOCAB
g = torch.Generator().manual_seed(0)
h = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
w = (torch.randn(V, K, generator=g) * 0.06).to(torch.bfloat16).cuda()
bias = torch.randn(V, generator=g).cuda()
tg = torch.stack([torch.randint(0, n, (M,), generator=g) for n in E.N_TOKENS], dim=1).to(torch.int32).cuda()
mk = (torch.rand(M, 8, generator=g) < 0.3).float().cuda()
den = mk.sum(0).clamp(min=1.0)
seg = (C.c_int * 8)(*E.N_TOKENS)
wts = (C.c_float * 8)(262, 134, 262, 134, 38, 135, 55, 260)
loss, cor = torch.zeros(8, device='cuda'), torch.zeros(8, device='cuda')
dl = torch.empty(M, V, dtype=torch.bfloat16, device='cuda')
logits = torch.empty(M, V, dtype=torch.float32, device='cuda')
plan = E.Plan(E.PB_BF16)
plan.gemm(h.data_ptr(), w.data_ptr(), logits.data_ptr(), M, V, K, K, K, V, bias=bias.data_ptr(), flags=L.PB_GEMM_OUT_F32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def fused():
    L.check(lib.pb_heads_ce_fused(P(h.data_ptr()), C.c_longlong(K), P(w.data_ptr()), P(bias.data_ptr()), P(tg.data_ptr()),
                                  P(mk.data_ptr()), P(den.data_ptr()), P(loss.data_ptr()), P(cor.data_ptr()), P(dl.data_ptr()),
                                  P(None), C.c_longlong(M), K, 8, seg, wts, C.c_float(1.0), L.stream_ptr()), 'fused')


def unfused():
    plan.run()
    L.check(lib.pb_heads_ce(P(logits.data_ptr()), P(tg.data_ptr()), P(mk.data_ptr()), P(den.data_ptr()), P(loss.data_ptr()),
                            P(cor.data_ptr()), P(dl.data_ptr()), P(None), C.c_longlong(M), 8, seg, wts, C.c_float(1.0),
                            E.PB_BF16, L.stream_ptr()), 'ce')


for name, fn in (('fused', fused), ('gemm+ce', unfused)):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print('%-8s median %.1f us  min %.1f us  (%.0f TFLOP/s of the %d x %d x %d product)' %
          (name, ts[len(ts) // 2], ts[0], 2.0 * M * V * K / ts[len(ts) // 2] / 1e6, M, V, K))

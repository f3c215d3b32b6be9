import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pianobart_b200 import _lib as L
lib = L.lib(); dev = 'cuda:0'
P = C.c_void_p
torch.manual_seed(0)
for dtype, code in ((torch.float32, 0), (torch.bfloat16, 1)):
    for M, d in ((37, 64), (50, 128), (300, 1024)):
        x = torch.randn(M, d, device=dev).to(dtype); g = torch.randn(d, device=dev); b = torch.randn(d, device=dev)
        y = torch.empty_like(x); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
        seed = torch.tensor([777], dtype=torch.int64, device=dev)
        site = L.DropSite(); site.seed = seed.data_ptr(); site.op = 5; site.thresh = int(0.9 * 2**32); site.scale = 1 / 0.9
        L.check(lib.pb_layernorm_fwd_drop(P(x.data_ptr()), P(g.data_ptr()), P(b.data_ptr()), P(y.data_ptr()), P(mean.data_ptr()), P(rstd.data_ptr()),
                                          C.c_longlong(M), d, C.c_float(1e-5), C.byref(site), code, L.stream_ptr()), 'ln')
        mk = torch.empty(M * d, dtype=torch.uint8, device=dev)
        L.check(lib.pb_dropout_mask(P(seed.data_ptr()), 5, site.thresh, P(mk.data_ptr()), C.c_longlong(M * d), L.stream_ptr()), 'mask')
        ref = torch.nn.functional.layer_norm(x.float(), (d,), g, b, 1e-5) * mk.view(M, d).float() / 0.9
        print(dtype, M, d, 'fwd err', (y.float() - ref).abs().max().item(), 'keep', mk.float().mean().item())
        # backward with in and out sites
        dy = torch.randn(M, d, device=dev).to(dtype); dx = torch.empty_like(x); dxd = torch.empty_like(x)
        dg = torch.zeros(d, device=dev); db = torch.zeros(d, device=dev); dbias = torch.zeros(d, device=dev)
        s_in = L.DropSite(); s_in.seed = seed.data_ptr(); s_in.op = 5; s_in.thresh = site.thresh; s_in.scale = 1 / 0.9
        s_out = L.DropSite(); s_out.seed = seed.data_ptr(); s_out.op = 9; s_out.thresh = site.thresh; s_out.scale = 1 / 0.9
        L.check(lib.pb_layernorm_bwd_drop(P(dy.data_ptr()), P(x.data_ptr()), P(g.data_ptr()), P(mean.data_ptr()), P(rstd.data_ptr()), P(dx.data_ptr()),
                                          P(dxd.data_ptr()), P(dg.data_ptr()), P(db.data_ptr()), P(dbias.data_ptr()), C.c_longlong(M), d,
                                          C.byref(s_in), C.byref(s_out), code, L.stream_ptr()), 'lnb')
        mk2 = torch.empty(M * d, dtype=torch.uint8, device=dev)
        L.check(lib.pb_dropout_mask(P(seed.data_ptr()), 9, site.thresh, P(mk2.data_ptr()), C.c_longlong(M * d), L.stream_ptr()), 'mask')
        xf = x.float().requires_grad_(True); gf = g.clone().requires_grad_(True); bf = b.clone().requires_grad_(True)
        out = torch.nn.functional.layer_norm(xf, (d,), gf, bf, 1e-5) * mk.view(M, d).float() / 0.9
        out.backward(dy.float())
        print('   bwd dx', (dx.float() - xf.grad).abs().max().item(), 'dxd', (dxd.float() - xf.grad * mk2.view(M, d).float() / 0.9).abs().max().item(),
              'dg', (dg - gf.grad).abs().max().item(), 'db', (db - bf.grad).abs().max().item(), 'dbias', (dbias - (xf.grad * mk2.view(M, d).float() / 0.9).sum(0)).abs().max().item())

# timing at the model's shape (M = 16 x 1024 rows, d = 1024, bf16), L2 flushed between launches by cycling 8 buffers
M, d = 16384, 1024
xs = [torch.randn(M, d, device=dev).bfloat16() for _ in range(8)]
ys = [torch.empty(M, d, device=dev, dtype=torch.bfloat16) for _ in range(8)]
g = torch.randn(d, device=dev); b = torch.randn(d, device=dev)
mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
dg = torch.zeros(d, device=dev); db = torch.zeros(d, device=dev); dbias = torch.zeros(d, device=dev)
dxd = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
seed = torch.tensor([777], dtype=torch.int64, device=dev)
s_in = L.DropSite(); s_in.seed = seed.data_ptr(); s_in.op = 5; s_in.thresh = int(0.9 * 2**32); s_in.scale = 1 / 0.9
def t(fn, n=40):
    for i in range(8): fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
f1 = lambda i: lib.pb_layernorm_fwd(P(xs[i % 8].data_ptr()), P(g.data_ptr()), P(b.data_ptr()), P(ys[i % 8].data_ptr()), P(mean.data_ptr()), P(rstd.data_ptr()), C.c_longlong(M), d, C.c_float(1e-5), 1, L.stream_ptr())
f2 = lambda i: lib.pb_layernorm_bwd_drop(P(ys[i % 8].data_ptr()), P(xs[i % 8].data_ptr()), P(g.data_ptr()), P(mean.data_ptr()), P(rstd.data_ptr()), P(ys[(i + 4) % 8].data_ptr()),
                                         P(dxd.data_ptr()), P(dg.data_ptr()), P(db.data_ptr()), P(dbias.data_ptr()), C.c_longlong(M), d, None, C.byref(s_in), 1, L.stream_ptr())
print('layernorm_fwd  M=16384 d=1024 bf16: %.1f us (64 MB moved: %.0f GB/s)' % (t(f1), 67.1e6 / t(f1) / 1e3))
print('layernorm_bwd (+dropout out-site) : %.1f us (134 MB moved: %.0f GB/s)' % (t(f2), 134.2e6 / t(f2) / 1e3))

"""Octuple embedding backward (scatter-add into the 8 tables): correctness vs torch index_add and timing at M = 16384."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pianobart_b200 import _lib as L
from pianobart_b200.vocab import build_octuple_vocab
lib = L.lib(); dev = 'cuda:0'; P = C.c_void_p
e2w, w2e = build_octuple_vocab()
ntok = [len(e2w[k]) for k in e2w]
off = [0]
for n in ntok[:-1]: off.append(off[-1] + n)
M = 16384
torch.manual_seed(0)
ids = torch.stack([torch.randint(0, n, (M,), device=dev) for n in ntok], 1).int().contiguous()
ids[:, 0] = (torch.arange(M, device=dev) // 64 % ntok[0]).int()     # skewed like real bars
dx = torch.randn(M, 2048, device=dev).bfloat16()
tab = torch.zeros(sum(ntok), 256, device=dev)
arr = (C.c_int * 8)(*ntok)
L.check(lib.pb_octuple_embed_bwd(P(ids.data_ptr()), 0, P(dx.data_ptr()), P(tab.data_ptr()), C.c_longlong(M), arr, C.c_float(16.0), 1, L.stream_ptr()))
ref = torch.zeros_like(tab)
for a in range(8):
    ref.index_add_(0, (ids[:, a].long() + off[a]), dx[:, a * 256:(a + 1) * 256].float() * 16.0)
print('max err', (tab - ref).abs().max().item(), 'ref max', ref.abs().max().item())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    lib.pb_octuple_embed_bwd(P(ids.data_ptr()), 0, P(dx.data_ptr()), P(tab.data_ptr()), C.c_longlong(M), arr, C.c_float(16.0), 1, L.stream_ptr())
e1.record(); torch.cuda.synchronize()
print('embed_bwd M=16384: %.1f us' % (e0.elapsed_time(e1) / 20 * 1e3))

import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pianobart_b200 import _lib as L
lib = L.lib(); dev = 'cuda:0'; torch.manual_seed(0)
B, H, S, hd = 16, 8, 1024, 128; d = H * hd
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
a.B, a.H, a.Sq, a.Sk, a.hd, a.causal, a.scale = B, H, S, S, hd, 0, hd ** -0.5
for _ in range(2):
    lib.pb_attn_fwd(C.byref(a), L.stream_ptr()); lib.pb_attn_bwd(C.byref(a), L.stream_ptr())
torch.cuda.synchronize()

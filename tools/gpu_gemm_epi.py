"""Epilogue cost study of the tcgen05 GEMM on the model's shapes (M = 16 x 1024 tokens): the same mainloop with the
epilogues the pretraining step uses.  Run on a B200: python tools/gpu_gemm_epi.py [reps]
(reps=1 under ncu: one launch per configuration, in the order printed)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pianobart_b200 import _lib as L

lib = L.lib()
dev = torch.device('cuda:0')
torch.manual_seed(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
M = 16384
seed = torch.tensor([12345], dtype=torch.int64, device=dev)


def case(tag, N, K, bias=False, drop=False, res=False, flags=0, aux=False, b_mn=0, block_n=0, cg=0):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    Bm = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    if b_mn:
        Bm = Bm.t().contiguous()
    Cout = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    bias_t = torch.randn(N, device=dev) if bias else None
    res_t = torch.randn(M, N, device=dev).bfloat16() if res else None
    aux_t = torch.randn(M, N, device=dev).bfloat16() if aux else None
    d = L.GemmDesc()
    d.a, d.b, d.c = A.data_ptr(), Bm.data_ptr(), Cout.data_ptr()
    d.bias = bias_t.data_ptr() if bias else None
    d.residual = res_t.data_ptr() if res else None
    d.M, d.N, d.K = M, N, K
    d.a_mn_major, d.b_mn_major = 0, b_mn
    d.lda, d.ldb, d.ldc, d.ldr = K, (N if b_mn else K), N, N
    d.batch_h = d.batch_b = 1
    d.alpha, d.flags, d.split_k = 1.0, flags, 1
    d.block_n, d.cta_group = block_n, cg
    if aux:
        d.aux, d.ldaux = aux_t.data_ptr(), N
    if drop:
        d.drop_seed, d.drop_op, d.drop_thresh, d.drop_scale = seed.data_ptr(), 3, int(0.9 * 2 ** 32), 1 / 0.9
    s = L.stream_ptr()
    for _ in range(2 if reps > 1 else 0):
        L.check(lib.pb_gemm_bf16(C.byref(d), s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        L.check(lib.pb_gemm_bf16(C.byref(d), s))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('%-34s N=%4d K=%4d  %7.1f us  %7.1f TFLOP/s' % (tag, N, K, ms * 1e3, 2.0 * M * N * K / ms / 1e9))


case('plain', 1024, 1024)
case('bias', 1024, 1024, bias=True)
case('bias+residual', 1024, 1024, bias=True, res=True)
case('out_proj: bias+dropout+residual', 1024, 1024, bias=True, drop=True, res=True)
case('qkv: bias', 3072, 1024, bias=True)
case('fc1: bias+gelu+preact', 4096, 1024, bias=True, flags=L.PB_GEMM_GELU | L.PB_GEMM_AUX_PREACT, aux=True)
case('dZ: mul dgelu(preact)', 4096, 1024, flags=L.PB_GEMM_MUL_DGELU, aux=True, b_mn=1)
case('fc1: bias+gelu+dgelu', 4096, 1024, bias=True, flags=L.PB_GEMM_GELU | L.PB_GEMM_AUX_DGELU, aux=True)
case('dZ: mul aux', 4096, 1024, flags=L.PB_GEMM_MUL_AUX, aux=True, b_mn=1)
case('fc2: bias+dropout+residual', 1024, 4096, bias=True, drop=True, res=True)
case('plain N=4096', 4096, 1024)
case('plain K=4096', 1024, 4096)
case('out_proj bn=128 (no pairs)', 1024, 1024, bias=True, drop=True, res=True, block_n=128)
case('plain bn=128 (no pairs)', 1024, 1024, block_n=128)
case('out_proj bn=256 no pairs', 1024, 1024, bias=True, drop=True, res=True, block_n=256, cg=1)
case('fc1 N=2048: bias+gelu+dgelu', 2048, 1024, bias=True, flags=L.PB_GEMM_GELU | L.PB_GEMM_AUX_DGELU, aux=True)
case('fc2 K=2048: bias+dropout+residual', 1024, 2048, bias=True, drop=True, res=True)
case('fc2 K=2048 bn=128', 1024, 2048, bias=True, drop=True, res=True, block_n=128)
case('plain N=2048', 2048, 1024)
case('bias N=2048', 2048, 1024, bias=True)
case('N=2048: bias+gelu (no second output)', 2048, 1024, bias=True, flags=L.PB_GEMM_GELU)
case('N=2048: bias+preact aux store only', 2048, 1024, bias=True, flags=L.PB_GEMM_AUX_PREACT, aux=True)

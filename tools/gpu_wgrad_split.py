"""Split-K sweep for the weight-gradient GEMMs (dW[n_out,n_in] += dY^T X over M = 16384 tokens, fp32 atomic accumulate)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pianobart_b200 import _lib as L
lib = L.lib(); dev = 'cuda:0'
M = 16384
for n_out, n_in in ((1024, 1024), (3072, 1024), (2048, 1024), (1024, 2048), (1280, 1024)):
    dy = (torch.randn(M, n_out, device=dev) * 0.1).bfloat16(); x = (torch.randn(M, n_in, device=dev) * 0.1).bfloat16()
    dw = torch.zeros(n_out, n_in, device=dev)
    res = []
    for split in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 18):
        d = L.GemmDesc()
        d.a, d.b, d.c = dy.data_ptr(), x.data_ptr(), dw.data_ptr()
        d.M, d.N, d.K = n_out, n_in, M
        d.a_mn_major = d.b_mn_major = 1
        d.lda, d.ldb, d.ldc = n_out, n_in, n_in
        d.batch_h = d.batch_b = 1
        d.alpha, d.flags, d.split_k, d.block_n = 1.0, L.PB_GEMM_OUT_F32 | L.PB_GEMM_ATOMIC_ACC, split, 256
        for _ in range(3): L.check(lib.pb_gemm_bf16(C.byref(d), L.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): L.check(lib.pb_gemm_bf16(C.byref(d), L.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        res.append((split, us))
    best = min(res, key=lambda r: r[1])
    print('%4d x %4d: ' % (n_out, n_in) + '  '.join('s%d:%.0f' % r for r in res) + '   best s%d %.0f us = %.0f TFLOP/s' % (best[0], best[1], 2.0 * M * n_out * n_in / best[1] / 1e6))

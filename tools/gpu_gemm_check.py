"""GPU bring-up check of the tcgen05 GEMM (all operand layouts, tails, batching, epilogues).
Run on a B200:  python tools/gpu_gemm_check.py     (torch is only the checker here)."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pianobart_b200 import _lib as L

lib = L.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)
fails = 0


def run(M, N, K, a_mn=0, b_mn=0, block_n=0, bias=False, gelu=False, res=False, out_f32=False, atomic=False,
        split_k=1, alpha=1.0, H=1, B=1, causal=0, tag="", cg=1):
    global fails
    # logical A [B,H,M,K], Bm [B,H,N,K]
    A = (torch.randn(B, H, M, K, device=dev) * 0.5).bfloat16()
    Bm = (torch.randn(B, H, N, K, device=dev) * 0.5).bfloat16()
    a_store = A.transpose(2, 3).contiguous() if a_mn else A.contiguous()
    b_store = Bm.transpose(2, 3).contiguous() if b_mn else Bm.contiguous()
    ref = torch.matmul(A.float(), Bm.float().transpose(2, 3)) * alpha
    bias_t = torch.randn(N, device=dev) if bias else None
    if bias: ref = ref + bias_t
    if gelu: ref = torch.nn.functional.gelu(ref)
    res_t = None
    if res:
        res_t = torch.randn(B, H, M, N, device=dev).bfloat16()
        ref = ref + res_t.float()
    if causal == 1:
        pass
    if causal == 2:
        # P.V semantics: only k <= m contribute when A is lower-triangular; make A lower triangular
        pass
    cdt = torch.float32 if out_f32 else torch.bfloat16
    Cout = torch.zeros(B, H, M, N, device=dev, dtype=cdt)
    if atomic:
        init = torch.randn(B, H, M, N, device=dev)
        Cout.copy_(init)
        ref = ref + init
    d = L.GemmDesc()
    d.a = a_store.data_ptr(); d.b = b_store.data_ptr(); d.c = Cout.data_ptr()
    d.bias = bias_t.data_ptr() if bias else None
    d.residual = res_t.data_ptr() if res else None
    d.M, d.N, d.K = M, N, K
    d.a_mn_major, d.b_mn_major = a_mn, b_mn
    d.lda = M if a_mn else K
    d.ldb = N if b_mn else K
    d.ldc = N; d.ldr = N
    d.batch_h, d.batch_b = H, B
    d.a_stride_h = M * K; d.a_stride_b = H * M * K
    d.b_stride_h = N * K; d.b_stride_b = H * N * K
    d.c_stride_h = M * N; d.c_stride_b = H * M * N
    d.r_stride_h = M * N; d.r_stride_b = H * M * N
    d.alpha = alpha
    d.flags = (L.PB_GEMM_OUT_F32 if out_f32 else 0) | (L.PB_GEMM_GELU if gelu else 0) | (L.PB_GEMM_ATOMIC_ACC if atomic else 0)
    d.split_k = split_k; d.causal = causal; d.block_n = block_n; d.cta_group = cg
    rc = lib.pb_gemm_bf16(C.byref(d), L.stream_ptr())
    if rc != 0:
        print("FAIL launch", tag, lib.pb_last_error().decode()); fails += 1; return
    torch.cuda.synchronize()
    out = Cout.float()
    if causal == 1:
        mask = torch.ones(M, N, device=dev, dtype=torch.bool).tril()
        # tiles fully above the diagonal are skipped (garbage allowed); compare the lower triangle only
        out = torch.where(mask, out, torch.zeros_like(out)); ref = torch.where(mask, ref, torch.zeros_like(ref))
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    rel = err / scale
    tol = 2e-2 if not out_f32 else 2e-3
    ok = rel < tol and torch.isfinite(out).all().item()
    print("%s %-44s M=%d N=%d K=%d a_mn=%d b_mn=%d bn=%d H=%d B=%d split=%d rel_err=%.3e" % (
        "ok  " if ok else "FAIL", tag, M, N, K, a_mn, b_mn, block_n, H, B, split_k, rel))
    if not ok: fails += 1


for a_mn in (0, 1):
    for b_mn in (0, 1):
        for bn in (128, 256):
            run(256, 512, 256, a_mn, b_mn, bn, out_f32=True, tag="layouts")
run(128, 256, 64, tag="single tile k=64")
run(1024, 1024, 1024, tag="square")
run(1000, 1280, 1024, bias=True, tag="M tail + bias (heads)")
run(384, 200, 136, bias=True, out_f32=True, tag="N tail, K tail (K%8==0)")
run(384, 200, 136, a_mn=1, b_mn=1, out_f32=True, tag="tails, MN-major (needs ld%8==0)")
run(512, 2048, 1024, bias=True, gelu=True, tag="fc1 bias+gelu")
run(512, 1024, 2048, bias=True, res=True, tag="fc2 bias+residual")
run(1024, 1024, 4096, a_mn=1, b_mn=1, out_f32=True, atomic=True, split_k=8, tag="dW split-K atomic")
run(256, 128, 128, H=8, B=2, alpha=0.088388, out_f32=True, tag="batched QK^T-like")
run(256, 128, 256, b_mn=1, H=8, B=2, tag="batched PV-like (B MN-major)")
run(1024, 1024, 128, H=2, B=1, out_f32=True, causal=1, tag="causal scores skip")
run(2048, 3072, 1024, bias=True, tag="QKV proj")
for a_mn in (0, 1):
    for b_mn in (0, 1):
        run(512, 512, 256, a_mn, b_mn, 256, out_f32=True, tag="cta_group::2 layouts", cg=2)
run(1000, 1280, 1024, bias=True, tag="cg2 M tail + bias", cg=2)
run(2048, 3072, 1024, bias=True, gelu=True, tag="cg2 bias+gelu", cg=2)
run(1024, 1024, 2048, bias=True, res=True, tag="cg2 bias+residual", cg=2)
run(1024, 1024, 4096, a_mn=1, b_mn=1, out_f32=True, atomic=True, split_k=8, tag="cg2 dW split-K atomic", cg=2)

# strided per-head views of a fused QKV activation [B,S,3,H,hd]
Bz, S, Hh, hd = 2, 256, 8, 128
qkv = (torch.randn(Bz, S, 3 * Hh * hd, device=dev) * 0.5).bfloat16()
q = qkv[:, :, 0:Hh * hd].view(Bz, S, Hh, hd).permute(0, 2, 1, 3)
k = qkv[:, :, Hh * hd:2 * Hh * hd].view(Bz, S, Hh, hd).permute(0, 2, 1, 3)
ref = torch.matmul(q.float(), k.float().transpose(2, 3))
out = torch.empty(Bz, Hh, S, S, device=dev)
d = L.GemmDesc()
d.a = qkv.data_ptr(); d.b = qkv.data_ptr() + Hh * hd * 2; d.c = out.data_ptr()
d.M, d.N, d.K = S, S, hd
d.lda = d.ldb = 3 * Hh * hd; d.ldc = S
d.batch_h, d.batch_b = Hh, Bz
d.a_stride_h = d.b_stride_h = hd; d.a_stride_b = d.b_stride_b = S * 3 * Hh * hd
d.c_stride_h = S * S; d.c_stride_b = Hh * S * S
d.alpha = 1.0; d.flags = L.PB_GEMM_OUT_F32; d.split_k = 1
rc = lib.pb_gemm_bf16(C.byref(d), L.stream_ptr()); torch.cuda.synchronize()
rel = ((out - ref).abs().max() / ref.abs().max()).item()
print("%s strided fused-QKV heads rel_err=%.3e rc=%d %s" % ("ok  " if rel < 2e-3 and rc == 0 else "FAIL", rel, rc, lib.pb_last_error().decode()))
if not (rel < 2e-3 and rc == 0): fails += 1

# throughput
def bench(M, N, K, a_mn=0, b_mn=0, bn=256, iters=20, split_k=1, atomic=False, cg=1):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    Bm = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    Cc = torch.zeros(M, N, device=dev, dtype=torch.float32 if atomic else torch.bfloat16)
    d = L.GemmDesc()
    d.a = A.data_ptr(); d.b = Bm.data_ptr(); d.c = Cc.data_ptr()
    d.M, d.N, d.K = M, N, K
    d.a_mn_major, d.b_mn_major = a_mn, b_mn
    d.lda = M if a_mn else K; d.ldb = N if b_mn else K; d.ldc = N
    d.alpha = 1.0; d.split_k = split_k; d.block_n = bn; d.cta_group = cg
    d.flags = (L.PB_GEMM_OUT_F32 | L.PB_GEMM_ATOMIC_ACC) if atomic else 0
    for _ in range(3): lib.pb_gemm_bf16(C.byref(d), L.stream_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): lib.pb_gemm_bf16(C.byref(d), L.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("bench cg=%d M=%d N=%d K=%d a_mn=%d b_mn=%d bn=%d split=%d: %.3f ms  %.1f TFLOP/s" % (cg, M, N, K, a_mn, b_mn, bn, split_k, ms, 2.0 * M * N * K / ms / 1e9))

bench(16384, 1024, 1024)
bench(16384, 1024, 1024, bn=128)
bench(16384, 3072, 1024)
bench(16384, 2048, 1024)
bench(16384, 1024, 2048)
bench(16384, 1024, 1024, b_mn=1)
bench(1024, 1024, 16384, a_mn=1, b_mn=1, split_k=5, atomic=True)
bench(2048, 1024, 16384, a_mn=1, b_mn=1, split_k=2, atomic=True)
bench(8192, 8192, 8192)
for cgv in (2,):
    bench(16384, 1024, 1024, cg=cgv)
    bench(16384, 3072, 1024, cg=cgv)
    bench(16384, 2048, 1024, cg=cgv)
    bench(16384, 1024, 2048, cg=cgv)
    bench(16384, 1024, 1024, b_mn=1, cg=cgv)
    bench(2048, 1024, 16384, a_mn=1, b_mn=1, split_k=2, atomic=True, cg=cgv)
    bench(8192, 8192, 8192, cg=cgv)
print("FAILS", fails)
sys.exit(1 if fails else 0)

// fp32 SIMT GEMM - the "fp32 parity mode" of the same graph the tcgen05 kernel runs in bf16.
// Same descriptor semantics as pb_gemm_bf16 (include/pianobart_b200.h).  It exists so that the
// whole path can be checked against the oracle at fp32 accuracy (north star: <= 1e-4 relative on
// the loss); it is not the production path and makes no performance claim.
#include "pb_internal.h"
#include "dropout.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float z) {
  return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * expf(-0.5f * z * z);
}

struct P {
  const float* a; const float* b; float* c; const float* bias; const float* residual; float* aux;
  int M, N, K;
  long long sa_m, sa_k, sb_n, sb_k, ldc, ldr, ldaux;
  int nh, nb;
  long long a_sh, a_sb, b_sh, b_sb, c_sh, c_sb, r_sh, r_sb;
  float alpha; int flags; int r_row_mod;
  pbdrop::Site drop;
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(const P p) {
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int batch = blockIdx.z;
  const int h = batch % p.nh, bb = batch / p.nh;
  const float* A = p.a + (long long)bb * p.a_sb + (long long)h * p.a_sh;
  const float* B = p.b + (long long)bb * p.b_sb + (long long)h * p.b_sh;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      int mm, kk;
      if (p.sa_k == 1) { kk = i % TK; mm = i / TK; } else { mm = i % TM; kk = i / TM; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < p.M && gk < p.K) ? A[(long long)gm * p.sa_m + (long long)gk * p.sa_k] : 0.f;
    }
    for (int i = threadIdx.x; i < TN * TK; i += 256) {
      int nn, kk;
      if (p.sb_k == 1) { kk = i % TK; nn = i / TK; } else { nn = i % TN; kk = i / TN; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < p.N && gk < p.K) ? B[(long long)gn * p.sb_n + (long long)gk * p.sb_k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float x = acc[i][j] * p.alpha;
      if (p.bias) x += p.bias[n];
      if (p.flags & PB_GEMM_AUX_PREACT) p.aux[(long long)m * p.ldaux + n] = x;
      if (p.flags & PB_GEMM_AUX_DGELU) p.aux[(long long)m * p.ldaux + n] = dgelu_erf(x);
      if (p.flags & PB_GEMM_GELU) x = gelu_erf(x);
      if (p.flags & PB_GEMM_MUL_DGELU) x *= dgelu_erf(p.aux[(long long)m * p.ldaux + n]);
      if (p.flags & PB_GEMM_MUL_AUX) x *= p.aux[(long long)m * p.ldaux + n];
      if (p.drop.seed) {
        const uint32_t key = pbdrop::site_key(*p.drop.seed, p.drop.op);
        x = pbdrop::keep(key, (unsigned long long)m * (unsigned long long)p.N + n, p.drop.thresh) ? x * p.drop.scale : 0.f;
      }
      if (p.residual) x += p.residual[(long long)bb * p.r_sb + (long long)h * p.r_sh + (long long)(p.r_row_mod > 0 ? m % p.r_row_mod : m) * p.ldr + n];
      float* c = p.c + (long long)bb * p.c_sb + (long long)h * p.c_sh + (long long)m * p.ldc + n;
      if (p.flags & PB_GEMM_ATOMIC_ACC) atomicAdd(c, x); else *c = x;
    }
  }
}

}  // namespace

extern "C" int pb_gemm_f32(const pb_gemm_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return pb_set_error("pb_gemm_f32: empty problem");
  P p;
  p.a = (const float*)d->a; p.b = (const float*)d->b; p.c = (float*)d->c;
  p.bias = d->bias; p.residual = (const float*)d->residual; p.aux = (float*)d->aux;
  p.M = d->M; p.N = d->N; p.K = d->K;
  if (d->a_mn_major) { p.sa_m = 1; p.sa_k = d->lda; } else { p.sa_m = d->lda; p.sa_k = 1; }
  if (d->b_mn_major) { p.sb_n = 1; p.sb_k = d->ldb; } else { p.sb_n = d->ldb; p.sb_k = 1; }
  p.ldc = d->ldc; p.ldr = d->ldr; p.ldaux = d->ldaux;
  p.nh = d->batch_h > 0 ? d->batch_h : 1; p.nb = d->batch_b > 0 ? d->batch_b : 1;
  p.a_sh = d->a_stride_h; p.a_sb = d->a_stride_b; p.b_sh = d->b_stride_h; p.b_sb = d->b_stride_b;
  p.c_sh = d->c_stride_h; p.c_sb = d->c_stride_b; p.r_sh = d->r_stride_h; p.r_sb = d->r_stride_b;
  p.alpha = d->alpha; p.flags = d->flags; p.r_row_mod = d->r_row_mod;
  p.drop.seed = d->drop_seed; p.drop.op = d->drop_op; p.drop.thresh = d->drop_thresh; p.drop.scale = d->drop_scale;
  if ((d->flags & (PB_GEMM_AUX_PREACT | PB_GEMM_MUL_DGELU | PB_GEMM_AUX_DGELU | PB_GEMM_MUL_AUX)) && d->aux == nullptr)
    return pb_set_error("pb_gemm_f32: aux epilogue without aux buffer");
  dim3 grid((d->N + TN - 1) / TN, (d->M + TM - 1) / TM, p.nh * p.nb);
  if (grid.y > 65535 || grid.z > 65535) return pb_set_error("pb_gemm_f32: grid too large");
  gemm_f32_kernel<<<grid, 256, 0, stream>>>(p);
  return pb_check_launch("gemm_f32_kernel");
}

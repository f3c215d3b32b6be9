// Classifier heads of the finetuning tasks (kernels K15 / K16) - everything around the
// backbone that is not a large GEMM:
//   SequenceClassification (model.py:128-143,195-218):  tanh(x Ws1^T) Ws2^T -> softmax over the SEQUENCE axis ->
//       attn^T x (r = 4 pooled rows) -> Dropout -> Linear(4096, 256) -> ReLU -> Linear(256, C)
//   TokenClassification (model.py:236-272):  Dropout -> Linear(1024, 256) -> ReLU -> Linear(256, C);  replacement decoder
//       front end Embeddings(C, 64) * 8 -> Linear(64, 1024) (PianoBart.py:9-16, finetune.py:194-198)
// The wide Linears go through pb_gemm_*; the kernels here are the rest: projections onto <= 16 outputs with the
// activation of the preceding layer folded into the operand load (no [M,256] ReLU / [M,128] tanh tensor is written),
// the sequence softmax, the attention pooling, dropout and the small-table gather.  All fp32 (the heads hold < 0.1 % of
// the step's arithmetic and both dtype modes share them); every backward accumulates (+=) parameter gradients.
#include "pb_internal.h"
#include "dropout.cuh"
#include <stdint.h>

#define PB_STREAM(s) reinterpret_cast<cudaStream_t>(s)

namespace {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };
constexpr int NMAX = 16;   // outputs of a "small" projection
constexpr int RMAX = 8;    // attention rows of the pooled head

__device__ __forceinline__ float act_f(float x, int act) {
  return act == ACT_RELU ? fmaxf(x, 0.f) : (act == ACT_TANH ? tanhf(x) : x);
}
// derivative of act at pre-activation x
__device__ __forceinline__ float dact_f(float x, int act) {
  if (act == ACT_RELU) return x > 0.f ? 1.f : 0.f;
  if (act == ACT_TANH) { const float t = tanhf(x); return 1.f - t * t; }
  return 1.f;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum / max over 256 threads (result broadcast); `red` holds 8 floats
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmaxf(t, red[i]);
  return t;
}
inline int blocks_for(long long work, int per_block, int per_sm = 8) {
  long long b = (work + per_block - 1) / per_block;
  const long long cap = (long long)pb_num_sms() * per_sm;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------- dropout (nn.Dropout(0.1) in front of both classifiers)
__global__ void __launch_bounds__(256) dropout_apply_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                            pbdrop::Site site) {
  pdl_entry();
  const uint32_t key = pbdrop::site_key(*site.seed, site.op);
  const long long npair = (n + 1) >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += (long long)gridDim.x * blockDim.x) {
    bool k0, k1;
    pbdrop::keep2(key, (unsigned long long)(2 * i), site.thresh, k0, k1);
    y[2 * i] = k0 ? x[2 * i] * site.scale : 0.f;
    if (2 * i + 1 < n) y[2 * i + 1] = k1 ? x[2 * i + 1] * site.scale : 0.f;
  }
}

// ---------------------------------------------------------------- projection onto N <= 16 outputs
// y[m, c] = sum_k act(x[m, k]) W[c, k] + b[c]: one warp per row, lanes stride over k
__global__ void __launch_bounds__(256) smalln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, long long M,
                                                         int N, int K, int act) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += nw) {
    float acc[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; ++c) acc[c] = 0.f;
    const float* xr = x + m * K;
    for (int k = lane; k < K; k += 32) {
      const float a = act_f(xr[k], act);
#pragma unroll
      for (int c = 0; c < NMAX; ++c)
        if (c < N) acc[c] = fmaf(a, w[(long long)c * K + k], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < NMAX; ++c) {
      if (c < N) {
        const float s = warp_sum(acc[c]);
        if (lane == 0) y[m * N + c] = s + (bias ? bias[c] : 0.f);
      }
    }
  }
}
// dx[m, k] = act'(x[m, k]) sum_c dy[m, c] W[c, k];  dW[c, k] += sum_m dy[m, c] act(x[m, k]);  db[c] += sum_m dy[m, c]
// CTA = SB_ROWS rows, thread = column k: the weight column and its gradient stay in registers over the rows
constexpr int SB_ROWS = 64;
__global__ void __launch_bounds__(256) smalln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ dy, float* __restrict__ dx,
                                                         float* __restrict__ dw, float* __restrict__ dbias, long long M,
                                                         int N, int K, int act) {
  pdl_entry();
  __shared__ float s_dy[SB_ROWS][NMAX];
  const long long m0 = (long long)blockIdx.x * SB_ROWS;
  const int rows = (int)((M - m0) < SB_ROWS ? (M - m0) : SB_ROWS);
  for (int i = threadIdx.x; i < SB_ROWS * NMAX; i += blockDim.x) {
    const int r = i / NMAX, c = i % NMAX;
    s_dy[r][c] = (r < rows && c < N) ? dy[(m0 + r) * N + c] : 0.f;
  }
  __syncthreads();
  if (dbias && threadIdx.x < N) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += s_dy[r][threadIdx.x];
    atomicAdd(dbias + threadIdx.x, s);
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float wc[NMAX], g[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; ++c) { wc[c] = c < N ? w[(long long)c * K + k] : 0.f; g[c] = 0.f; }
    for (int r = 0; r < rows; ++r) {
      const float xv = x[(m0 + r) * K + k];
      const float a = act_f(xv, act);
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < NMAX; ++c) {
        const float t = s_dy[r][c];
        d = fmaf(t, wc[c], d);
        g[c] = fmaf(t, a, g[c]);
      }
      if (dx) dx[(m0 + r) * K + k] = d * dact_f(xv, act);
    }
    if (dw) {
#pragma unroll
      for (int c = 0; c < NMAX; ++c)
        if (c < N) atomicAdd(dw + (long long)c * K + k, g[c]);
    }
  }
}

// ---------------------------------------------------------------- softmax over the sequence axis (model.py:142, dim=1)
// a, p: [B, S, R]; CTA = one sequence, every (b, r) column normalised over s.  Padding rows are NOT masked (reference).
__global__ void __launch_bounds__(256) seq_softmax_fwd_kernel(const float* __restrict__ a, float* __restrict__ p, int S, int R) {
  pdl_entry();
  __shared__ float red[8];
  const float* ab = a + (long long)blockIdx.x * S * R;
  float* pb = p + (long long)blockIdx.x * S * R;
  for (int r = 0; r < R; ++r) {
    float mx = -INFINITY;
    for (int s = threadIdx.x; s < S; s += blockDim.x) mx = fmaxf(mx, ab[s * R + r]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int s = threadIdx.x; s < S; s += blockDim.x) sum += __expf(ab[s * R + r] - mx);
    sum = block_sum(sum, red);
    const float inv = 1.f / sum;
    for (int s = threadIdx.x; s < S; s += blockDim.x) pb[s * R + r] = __expf(ab[s * R + r] - mx) * inv;
  }
}
// da = p * (dp - sum_s p dp)
__global__ void __launch_bounds__(256) seq_softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                              float* __restrict__ da, int S, int R) {
  pdl_entry();
  __shared__ float red[8];
  const long long o = (long long)blockIdx.x * S * R;
  for (int r = 0; r < R; ++r) {
    float dot = 0.f;
    for (int s = threadIdx.x; s < S; s += blockDim.x) dot = fmaf(p[o + s * R + r], dp[o + s * R + r], dot);
    dot = block_sum(dot, red);
    for (int s = threadIdx.x; s < S; s += blockDim.x) da[o + s * R + r] = p[o + s * R + r] * (dp[o + s * R + r] - dot);
  }
}

// ---------------------------------------------------------------- attention pooling  m[b, r, :] = sum_s p[b, s, r] x[b, s, :]
// grid (D / 256, S chunks, B); thread = feature column; partial sums over a chunk of the sequence, fp32 atomics into m
constexpr int POOL_CHUNK = 128;
__global__ void __launch_bounds__(256) attn_pool_fwd_kernel(const float* __restrict__ p, const float* __restrict__ x,
                                                            float* __restrict__ m, int S, int R, int D) {
  pdl_entry();
  __shared__ float s_p[POOL_CHUNK][RMAX];
  const int b = blockIdx.z, s0 = blockIdx.y * POOL_CHUNK;
  const int ns = min(POOL_CHUNK, S - s0);
  for (int i = threadIdx.x; i < POOL_CHUNK * RMAX; i += blockDim.x) {
    const int s = i / RMAX, r = i % RMAX;
    s_p[s][r] = (s < ns && r < R) ? p[((long long)b * S + s0 + s) * R + r] : 0.f;
  }
  __syncthreads();
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float acc[RMAX];
#pragma unroll
  for (int r = 0; r < RMAX; ++r) acc[r] = 0.f;
  const float* xb = x + ((long long)b * S + s0) * D + d;
  for (int s = 0; s < ns; ++s) {
    const float xv = xb[(long long)s * D];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[r] = fmaf(s_p[s][r], xv, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < RMAX; ++r)
    if (r < R) atomicAdd(m + ((long long)b * R + r) * D + d, acc[r]);
}
// dx[b, s, :] = sum_r p[b, s, r] dm[b, r, :];  dp[b, s, r] = x[b, s, :] . dm[b, r, :]   (one warp per (b, s) row)
__global__ void __launch_bounds__(256) attn_pool_bwd_kernel(const float* __restrict__ p, const float* __restrict__ x,
                                                            const float* __restrict__ dm, float* __restrict__ dx,
                                                            float* __restrict__ dp, long long BS, int S, int R, int D) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < BS; row += nw) {
    const int b = (int)(row / S);
    float pr[RMAX], acc[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) { pr[r] = r < R ? p[row * R + r] : 0.f; acc[r] = 0.f; }
    const float* dmb = dm + (long long)b * R * D;
    for (int d = lane; d < D; d += 32) {
      const float xv = x[row * D + d];
      float g = 0.f;
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R) {
          const float t = dmb[(long long)r * D + d];
          g = fmaf(pr[r], t, g);
          acc[r] = fmaf(xv, t, acc[r]);
        }
      }
      dx[row * D + d] = g;
    }
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < R) {
        const float s = warp_sum(acc[r]);
        if (lane == 0) dp[row * R + r] = s;
      }
    }
  }
}

// ---------------------------------------------------------------- small-table lookup (Embeddings.forward, PianoBart.py:15-16)
__global__ void __launch_bounds__(256) rows_gather_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                                          float* __restrict__ out, long long M, int n_rows, int d,
                                                          float scale, int* __restrict__ err) {
  pdl_entry();
  const long long total = M * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / d;
    const int c = (int)(i % d);
    const long long id = ids[m];
    if (id < 0 || id >= n_rows) {
      if (err) atomicExch(err, 1);
      out[i] = 0.f;
    } else {
      out[i] = table[id * d + c] * scale;
    }
  }
}
// dtable[ids[m], :] += scale * dout[m, :]: the table is tiny (class_num x 64), so a CTA first accumulates its rows in
// shared memory and then issues one global atomic per table element
__global__ void __launch_bounds__(256) rows_scatter_kernel(const long long* __restrict__ ids, const float* __restrict__ dout,
                                                           float* __restrict__ dtable, long long M, int n_rows, int d,
                                                           float scale, int rows_per_cta) {
  pdl_entry();
  extern __shared__ float s_tab[];
  const int nt = n_rows * d;
  for (int i = threadIdx.x; i < nt; i += blockDim.x) s_tab[i] = 0.f;
  __syncthreads();
  const long long m0 = (long long)blockIdx.x * rows_per_cta;
  const long long m1 = (m0 + rows_per_cta < M) ? m0 + rows_per_cta : M;
  for (long long i = m0 * d + threadIdx.x; i < m1 * d; i += blockDim.x) {
    const long long id = ids[i / d];
    if (id >= 0 && id < n_rows) atomicAdd(&s_tab[id * d + (int)(i % d)], dout[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nt; i += blockDim.x)
    if (s_tab[i] != 0.f) atomicAdd(dtable + i, s_tab[i] * scale);
}

}  // namespace

extern "C" int pb_dropout_apply(const float* x, float* y, long long n, const pb_drop_site* site, void* stream) {
  if (!site || !site->seed) return pb_set_error("dropout_apply: site with a device seed required");
  pbdrop::Site st{site->seed, site->op, site->thresh, site->scale};
  PB_LAUNCH((dropout_apply_kernel), blocks_for((n + 1) / 2, 256 * 4), 256, 0, PB_STREAM(stream), x, y, n, st);
  return pb_check_launch("dropout_apply");
}

extern "C" int pb_smalln_linear_fwd(const float* x, const float* w, const float* bias, float* y, long long M, int N, int K,
                                    int act, void* stream) {
  if (N < 1 || N > NMAX) return pb_set_error("smalln_linear: 1 <= N <= 16");
  if (act < 0 || act > 2) return pb_set_error("smalln_linear: act must be 0 (none), 1 (relu) or 2 (tanh)");
  PB_LAUNCH((smalln_fwd_kernel), blocks_for(M, 8), 256, 0, PB_STREAM(stream), x, w, bias, y, M, N, K, act);
  return pb_check_launch("smalln_linear_fwd");
}

extern "C" int pb_smalln_linear_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias,
                                    long long M, int N, int K, int act, void* stream) {
  if (N < 1 || N > NMAX) return pb_set_error("smalln_linear: 1 <= N <= 16");
  if (act < 0 || act > 2) return pb_set_error("smalln_linear: act must be 0 (none), 1 (relu) or 2 (tanh)");
  const long long grid = (M + SB_ROWS - 1) / SB_ROWS;
  PB_LAUNCH((smalln_bwd_kernel), (unsigned)grid, 256, 0, PB_STREAM(stream), x, w, dy, dx, dw, dbias, M, N, K, act);
  return pb_check_launch("smalln_linear_bwd");
}

extern "C" int pb_seq_softmax_fwd(const float* a, float* p, int B, int S, int R, void* stream) {
  if (R < 1 || R > RMAX) return pb_set_error("seq_softmax: 1 <= R <= 8");
  PB_LAUNCH((seq_softmax_fwd_kernel), B, 256, 0, PB_STREAM(stream), a, p, S, R);
  return pb_check_launch("seq_softmax_fwd");
}

extern "C" int pb_seq_softmax_bwd(const float* p, const float* dp, float* da, int B, int S, int R, void* stream) {
  if (R < 1 || R > RMAX) return pb_set_error("seq_softmax: 1 <= R <= 8");
  PB_LAUNCH((seq_softmax_bwd_kernel), B, 256, 0, PB_STREAM(stream), p, dp, da, S, R);
  return pb_check_launch("seq_softmax_bwd");
}

extern "C" int pb_attn_pool_fwd(const float* p, const float* x, float* m, int B, int S, int R, int D, void* stream) {
  if (R < 1 || R > RMAX) return pb_set_error("attn_pool: 1 <= R <= 8");
  cudaError_t e = cudaMemsetAsync(m, 0, sizeof(float) * (size_t)B * R * D, PB_STREAM(stream));
  if (e != cudaSuccess) return pb_set_cuda_error("attn_pool: memset", e);
  dim3 grid((D + 255) / 256, (S + POOL_CHUNK - 1) / POOL_CHUNK, B);
  PB_LAUNCH((attn_pool_fwd_kernel), grid, 256, 0, PB_STREAM(stream), p, x, m, S, R, D);
  return pb_check_launch("attn_pool_fwd");
}

extern "C" int pb_attn_pool_bwd(const float* p, const float* x, const float* dm, float* dx, float* dp, int B, int S, int R,
                                int D, void* stream) {
  if (R < 1 || R > RMAX) return pb_set_error("attn_pool: 1 <= R <= 8");
  const long long BS = (long long)B * S;
  PB_LAUNCH((attn_pool_bwd_kernel), blocks_for(BS, 8), 256, 0, PB_STREAM(stream), p, x, dm, dx, dp, BS, S, R, D);
  return pb_check_launch("attn_pool_bwd");
}

extern "C" int pb_rows_gather(const long long* ids, const float* table, float* out, long long M, int n_rows, int d,
                              float scale, int* err_flag, void* stream) {
  PB_LAUNCH((rows_gather_kernel), blocks_for(M * d, 256 * 4), 256, 0, PB_STREAM(stream), ids, table, out, M, n_rows, d, scale, err_flag);
  return pb_check_launch("rows_gather");
}

extern "C" int pb_rows_scatter_add(const long long* ids, const float* dout, float* dtable, long long M, int n_rows, int d,
                                   float scale, void* stream) {
  const size_t smem = sizeof(float) * (size_t)n_rows * d;
  if (smem > 48 * 1024) return pb_set_error("rows_scatter_add: table larger than 48 KB (n_rows * d <= 12288)");
  const int rows_per_cta = 256;
  const long long grid = (M + rows_per_cta - 1) / rows_per_cta;
  PB_LAUNCH((rows_scatter_kernel), (unsigned)grid, 256, smem, PB_STREAM(stream), ids, dout, dtable, M, n_rows, d, scale, rows_per_cta);
  return pb_check_launch("rows_scatter_add");
}

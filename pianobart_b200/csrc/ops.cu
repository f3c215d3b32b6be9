// HBM-bound kernels of the PianoBART path (everything that is not a dense contraction):
// Octuple front end gather, LayerNorm, masked softmax, bias-gradient column sums, the fused
// 8-head masked cross-entropy (+argmax accuracy, +dlogits), gradient-norm, HF-semantics AdamW,
// dtype casts.  All kernels are templated on the activation type T (float = fp32 parity mode,
// __nv_bfloat16 = production mode), use 16-byte vector accesses and fp32 arithmetic.
#include "pb_internal.h"
#include "dropout.cuh"
#include <cuda_bf16.h>
#include <stdint.h>

namespace {

typedef __nv_bfloat16 bf16;

template <typename T> struct Pack;
template <> struct Pack<float> { static constexpr int N = 4; };
template <> struct Pack<bf16> { static constexpr int N = 8; };

__device__ __forceinline__ void load_pack(const float* p, float (&f)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void load_pack(const bf16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void store_pack(float* p, const float (&f)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
__device__ __forceinline__ void store_pack(bf16* p, const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int grid_for(long long work, int per_block, int max_blocks_per_sm = 16) {
  long long b = (work + per_block - 1) / per_block;
  long long cap = (long long)pb_num_sms() * max_blocks_per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ Octuple front end
// out[m, 256*i + c] = table[row_off[i] + ids[m,i], c]   (table already scaled by sqrt(256)=16)
struct EmbedMeta { int row_off[8]; int n_tok[8]; };

template <typename T, typename I>
__global__ void __launch_bounds__(256) octuple_embed_fwd_kernel(const I* __restrict__ ids, const T* __restrict__ table,
                                                                T* __restrict__ out, long long M, EmbedMeta meta,
                                                                int* __restrict__ err) {
  pdl_entry();
  constexpr int N = Pack<T>::N;
  constexpr int PACKS = 2048 / N;
  for (long long m = blockIdx.x; m < M; m += gridDim.x) {
    for (int pk = threadIdx.x; pk < PACKS; pk += blockDim.x) {
      const int col = pk * N;
      const int attr = col >> 8;
      long long id = (long long)ids[m * 8 + attr];
      if (id < 0 || id >= meta.n_tok[attr]) { if (err) atomicExch(err, 1); id = 0; }
      float f[N];
      load_pack(table + ((long long)(meta.row_off[attr] + id) << 8) + (col & 255), f);
      store_pack(out + m * 2048 + col, f);
    }
  }
}

// dtable[row_off[i] + ids[m,i], c] += scale * dx[m, 256*i + c]
// The Octuple vocabulary is tiny (<= 1280 rows over the 8 tables), so a scatter with global atomics is one long collision
// (33.5 M atomics on a few hundred hot rows per step: 266 us).  Instead each CTA owns a 32-column slice of every table in
// shared memory (1280 x 32 fp32 = 160 KB), accumulates a chunk of tokens into it with shared-memory atomics and flushes
// the non-zero entries once: 144 CTAs x <= 41 k global atomics.
constexpr int EB_COLS = 32;
template <typename T, typename I>
__global__ void __launch_bounds__(256) octuple_embed_bwd_kernel(const I* __restrict__ ids, const T* __restrict__ dx,
                                                                float* __restrict__ dtable, long long M, EmbedMeta meta,
                                                                float scale, int total_rows, long long tokens_per_cta) {
  pdl_entry();
  extern __shared__ float eb_acc[];          // [total_rows][32]; column c of a row of attribute a is stored at (c + a) & 31
  const int slice = blockIdx.x;              // columns [slice*32, slice*32 + 32) of every 256-wide attribute block
  const int n = total_rows * EB_COLS;
  for (int i = threadIdx.x; i < n; i += blockDim.x) eb_acc[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int attr = lane >> 2, piece = lane & 3;
  const long long m0 = (long long)blockIdx.y * tokens_per_cta;
  const long long m1 = m0 + tokens_per_cta < M ? m0 + tokens_per_cta : M;
  constexpr int U = 8;                       // tokens in flight per warp: both dependent loads (id, then its dx piece) are batched
  for (long long mb = m0 + (long long)warp * U; mb < m1; mb += (long long)(blockDim.x >> 5) * U) {
    long long id[U];
#pragma unroll
    for (int u = 0; u < U; ++u) id[u] = (mb + u < m1) ? (long long)ids[(mb + u) * 8 + attr] : -1;
    float f[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (id[u] < 0 || id[u] >= meta.n_tok[attr]) continue;
      const T* src = dx + (mb + u) * 2048 + attr * 256 + slice * EB_COLS + piece * 8;
      if constexpr (sizeof(T) == 2) {
        load_pack(src, f[u]);
      } else {
        float a[4], b4[4];
        load_pack(src, a); load_pack(src + 4, b4);
#pragma unroll
        for (int j = 0; j < 4; ++j) { f[u][j] = a[j]; f[u][4 + j] = b4[j]; }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (id[u] < 0 || id[u] >= meta.n_tok[attr]) continue;
      float* row = eb_acc + (meta.row_off[attr] + (int)id[u]) * EB_COLS;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(row + ((piece * 8 + j + attr) & 31), f[u][j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = eb_acc[i];
    if (v == 0.f) continue;
    const int r = i >> 5;
    int a = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) a += (r >= meta.row_off[k]) ? 1 : 0;
    const int c = ((i & 31) - a) & 31;
    atomicAdd(dtable + ((long long)r << 8) + slice * EB_COLS + c, v * scale);
  }
}

// ------------------------------------------------------------------ LayerNorm (one warp per row)
// MAXP = 16-byte packs per lane (d <= MAXP * 32 * pack width); the row lives in registers.
template <typename T> struct RawPack;
template <> struct RawPack<float> { typedef float4 type; };
template <> struct RawPack<bf16> { typedef uint4 type; };
__device__ __forceinline__ void unpack_raw(const float4& r, float (&f)[4]) { f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w; }
__device__ __forceinline__ void unpack_raw(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// gamma / beta are staged in shared memory once per block and the 16-byte loads of the warp's next row are in flight while
// the current row is reduced; the optional output dropout site costs one hash per two elements (dropout.cuh).
template <typename T, int MAXP>
__global__ void __launch_bounds__(128) layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, T* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            long long M, int d, float eps, pbdrop::Site drop) {
  pdl_entry();
  constexpr int N = Pack<T>::N;
  typedef typename RawPack<T>::type Raw;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const uint32_t dkey = drop.seed ? pbdrop::site_key(*drop.seed, drop.op) : 0u;
  __shared__ __align__(16) float s_ga[MAXP * 32 * N], s_be[MAXP * 32 * N];   // gamma / beta (registers would halve the occupancy)
  for (int c = threadIdx.x; c < MAXP * 32 * N; c += blockDim.x) {
    s_ga[c] = c < d ? __ldg(gamma + c) : 0.f;
    s_be[c] = c < d ? __ldg(beta + c) : 0.f;
  }
  __syncthreads();
  Raw nx[MAXP];
  auto fetch = [&](long long row) {
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) nx[k] = *reinterpret_cast<const Raw*>(x + row * d + c);
    }
  };
  long long row = warp_global;
  if (row < M) fetch(row);
  for (; row < M; row += nwarps) {
    float v[MAXP][N];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) {
        unpack_raw(nx[k], v[k]);
#pragma unroll
        for (int j = 0; j < N; ++j) s += v[k][j];
      }
    }
    if (row + nwarps < M) fetch(row + nwarps);
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) {
#pragma unroll
        for (int j = 0; j < N; ++j) { const float t = v[k][j] - mean; q += t * t; }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    T* yr = y + row * d;
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) {
        float o[N], ga[N], be[N];
#pragma unroll
        for (int j = 0; j < N; j += 4) {
          *reinterpret_cast<float4*>(&ga[j]) = *reinterpret_cast<const float4*>(&s_ga[c + j]);
          *reinterpret_cast<float4*>(&be[j]) = *reinterpret_cast<const float4*>(&s_be[c + j]);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = (v[k][j] - mean) * rstd * ga[j] + be[j];
        if (drop.seed) {
          const uint32_t bits = pbdrop::keep_bits<N>(dkey, (unsigned long long)row * d + c, drop.thresh);
#pragma unroll
          for (int j = 0; j < N; ++j) o[j] = ((bits >> j) & 1u) ? o[j] * drop.scale : 0.f;
        }
        store_pack(yr + c, o);
      }
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma;  dgamma += dy*xhat, dbeta += dy,
// dbias (optional) += dx  - the bias gradient of the Linear whose output (+ residual) fed this LayerNorm.
constexpr int LNB_WARPS = 8;

// One warp per row, rows software-pipelined: the 16-byte loads of row r+1 are in flight while row r is reduced
// (8 warps per SM because of the per-column accumulators in registers).  gamma is staged in shared memory; the keep
// bits of the input dropout site are generated once per row and reused by both passes.
template <typename T, int MAXP>
__global__ void __launch_bounds__(LNB_WARPS * 32) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, T* __restrict__ dx,
                                                            T* __restrict__ dx_drop, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, float* __restrict__ dbias,
                                                            long long M, int d, pbdrop::Site din, pbdrop::Site dout) {
  pdl_entry();
  constexpr int N = Pack<T>::N;
  typedef typename RawPack<T>::type Raw;
  extern __shared__ float sh_red[];   // LNB_WARPS * d floats (gamma occupies the first d until the final reduction)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t kin = din.seed ? pbdrop::site_key(*din.seed, din.op) : 0u;
  const uint32_t kout = dout.seed ? pbdrop::site_key(*dout.seed, dout.op) : 0u;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (int c = threadIdx.x; c < d; c += blockDim.x) sh_red[c] = gamma[c];
  __syncthreads();
  float ag[MAXP][N], ab[MAXP][N], ax[MAXP][N];
#pragma unroll
  for (int k = 0; k < MAXP; ++k)
#pragma unroll
    for (int j = 0; j < N; ++j) { ag[k][j] = 0.f; ab[k][j] = 0.f; ax[k][j] = 0.f; }
  Raw rx[MAXP], rdy[MAXP];
  float mean = 0.f, rstd = 0.f;
  auto fetch = [&](long long row) {
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) {
        rx[k] = *reinterpret_cast<const Raw*>(x + row * d + c);
        rdy[k] = *reinterpret_cast<const Raw*>(dy + row * d + c);
      }
    }
    mean = mean_in[row];
    rstd = rstd_in[row];
  };
  long long row = warp_global;
  if (row < M) fetch(row);
  for (; row < M; row += nwarps) {
    float xh[MAXP][N], g[MAXP][N];     // normalised input and dy * gamma of this lane's columns
    const float cm = mean, cr = rstd;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) {
        float dv[N], gm[N];
        unpack_raw(rx[k], xh[k]);
        unpack_raw(rdy[k], dv);
#pragma unroll
        for (int j = 0; j < N; j += 4) *reinterpret_cast<float4*>(&gm[j]) = *reinterpret_cast<const float4*>(&sh_red[c + j]);
        if (din.seed) {
          const uint32_t bits = pbdrop::keep_bits<N>(kin, (unsigned long long)row * d + c, din.thresh);
#pragma unroll
          for (int j = 0; j < N; ++j) dv[j] = ((bits >> j) & 1u) ? dv[j] * din.scale : 0.f;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
          xh[k][j] = (xh[k][j] - cm) * cr;
          g[k][j] = dv[j] * gm[j];
          s1 += g[k][j];
          s2 = fmaf(g[k][j], xh[k][j], s2);
          ag[k][j] = fmaf(dv[j], xh[k][j], ag[k][j]);
          ab[k][j] += dv[j];
        }
      }
    }
    if (row + nwarps < M) fetch(row + nwarps);
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) {
        float o[N];
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = cr * (g[k][j] - s1 - xh[k][j] * s2);
        store_pack(dx + row * d + c, o);
        if (dout.seed) {  // gradient of the dropped-out Linear output that (plus the residual) fed this LayerNorm
          const uint32_t bits = pbdrop::keep_bits<N>(kout, (unsigned long long)row * d + c, dout.thresh);
#pragma unroll
          for (int j = 0; j < N; ++j) o[j] = ((bits >> j) & 1u) ? o[j] * dout.scale : 0.f;
          store_pack(dx_drop + row * d + c, o);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) ax[k][j] += o[j];     // the Linear's bias sits inside the dropout: its gradient sums the dropped dx
      }
    }
  }
  // block reduction of the per-warp column partials: every warp publishes its d partials, then all threads
  // sum the LNB_WARPS copies of their columns and issue one atomic per column per block
  for (int pass = 0; pass < 3; ++pass) {
    float* out = pass == 0 ? dgamma : (pass == 1 ? dbeta : dbias);
    if (out == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) {
#pragma unroll
        for (int j = 0; j < N; ++j) sh_red[warp * d + c + j] = pass == 0 ? ag[k][j] : (pass == 1 ? ab[k][j] : ax[k][j]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < LNB_WARPS; ++w) t += sh_red[w * d + c];
      atomicAdd(out + c, t);
    }
  }
}

// ------------------------------------------------------------------ masked softmax (one warp per row)
// scores fp32 [B,H,Sq,Sk] (already scaled) -> probs T.  key_keep uint8 [B,Sk] (may be NULL), causal: j <= i.
constexpr int SM_MAX = 32;  // Sk <= 1024

template <typename T>
__global__ void __launch_bounds__(128) softmax_fwd_kernel(const float* __restrict__ s, T* __restrict__ p,
                                                          const uint8_t* __restrict__ key_keep, int B, int H, int Sq,
                                                          int Sk, int causal) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * H * Sq;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp_global; row < rows; row += nwarps) {
    const int i = (int)(row % Sq);
    const int b = (int)(row / ((long long)H * Sq));
    const float* sr = s + row * Sk;
    const uint8_t* kk = key_keep ? key_keep + (long long)b * Sk : nullptr;
    const int jmax = causal ? min(Sk, i + 1) : Sk;
    float v[SM_MAX];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < SM_MAX; ++k) {
      const int j = k * 32 + lane;
      float x = -INFINITY;
      if (j < jmax && (!kk || kk[j])) x = sr[j];
      v[k] = x;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < SM_MAX; ++k) {
      const float e = (v[k] == -INFINITY) ? 0.f : __expf(v[k] - mx);
      v[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    T* pr = p + row * Sk;
#pragma unroll
    for (int k = 0; k < SM_MAX; ++k) {
      const int j = k * 32 + lane;
      if (j < Sk) pr[j] = from_f<T>(v[k] * inv);
    }
  }
}

// ds = p * (dp - sum_j p_j dp_j)  (positions outside the mask are forced to 0; dp there may be garbage)
template <typename T>
__global__ void __launch_bounds__(128) softmax_bwd_kernel(const T* __restrict__ p, const float* __restrict__ dp,
                                                          T* __restrict__ ds, const uint8_t* __restrict__ key_keep,
                                                          int B, int H, int Sq, int Sk, int causal) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * H * Sq;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp_global; row < rows; row += nwarps) {
    const int i = (int)(row % Sq);
    const int b = (int)(row / ((long long)H * Sq));
    const uint8_t* kk = key_keep ? key_keep + (long long)b * Sk : nullptr;
    const int jmax = causal ? min(Sk, i + 1) : Sk;
    float pv[SM_MAX], dv[SM_MAX];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < SM_MAX; ++k) {
      const int j = k * 32 + lane;
      float a = 0.f, g = 0.f;
      if (j < jmax && (!kk || kk[j])) { a = to_f(p[row * Sk + j]); g = dp[row * Sk + j]; }
      pv[k] = a; dv[k] = g;
      dot += a * g;
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < SM_MAX; ++k) {
      const int j = k * 32 + lane;
      if (j < Sk) ds[row * Sk + j] = from_f<T>(pv[k] * (dv[k] - dot));
    }
  }
}

// ------------------------------------------------------------------ column sums (bias gradients)
// block = 32 column packs x 8 row lanes; every thread streams 16-byte packs down its rows, the 8 row lanes are
// reduced through shared memory and one atomic per column per block is issued.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long M, int N,
                                                     long long ld, int rows_per_block) {
  pdl_entry();
  constexpr int PN = Pack<T>::N;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cx) * PN;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float acc[PN];
#pragma unroll
  for (int j = 0; j < PN; ++j) acc[j] = 0.f;
  if (c < N) {
    for (long long r = r0 + ry; r < r1; r += 8) {
      float f[PN];
      load_pack(x + r * ld + c, f);
#pragma unroll
      for (int j = 0; j < PN; ++j) acc[j] += f[j];
    }
  }
  __shared__ float sh[8][32 * PN + 1];
#pragma unroll
  for (int j = 0; j < PN; ++j) sh[ry][cx * PN + j] = acc[j];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * PN; i += 256) {
    const int col = blockIdx.x * 32 * PN + i;
    if (col < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sh[w][i];
      atomicAdd(out + col, t);
    }
  }
}

// ------------------------------------------------------------------ fused multi-head masked CE
struct SegMeta { int nseg; int off[17]; float w[16]; };

// logits fp32 [M, V] (V = off[nseg]); targets int32 [M, nseg]; mask float [M, nseg];
// coef[s] = w[s] / (sum_w * den[s]) computed on the fly from den (global mask sums).
// outputs: loss_num[s] += (lse - logit_t) * mask ; correct[s] += (argmax == t) * mask ;
//          dlogits[m, j] = (softmax_j - [j==t]) * mask * coef[s] * grad_scale      (if dlogits != NULL)
template <typename T>
__global__ void __launch_bounds__(128) heads_ce_kernel(const float* __restrict__ logits, const int* __restrict__ targets,
                                                       const float* __restrict__ mask, const float* __restrict__ den,
                                                       float* __restrict__ loss_num, float* __restrict__ correct,
                                                       T* __restrict__ dlogits, int* __restrict__ argmax_out, long long M,
                                                       SegMeta meta, float sum_w, float grad_scale) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int V = meta.off[meta.nseg];
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float acc_loss[16], acc_cor[16];
#pragma unroll
  for (int s = 0; s < 16; ++s) { acc_loss[s] = 0.f; acc_cor[s] = 0.f; }
  for (long long row = warp_global; row < M; row += nwarps) {
    const float* lr = logits + row * V;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      if (s < meta.nseg) {
        const int o = meta.off[s], n = meta.off[s + 1] - meta.off[s];
        const int t = targets[row * meta.nseg + s];
        const float mk = mask[row * meta.nseg + s];
        float mx = -INFINITY;
        int am = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
          const float x = lr[o + j];
          if (x > mx) { mx = x; am = j; }
        }
        // warp argmax with lowest-index tie break (numpy argmax semantics, pretrain.py:165)
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
          const float omx = __shfl_xor_sync(0xffffffffu, mx, sft);
          const int oam = __shfl_xor_sync(0xffffffffu, am, sft);
          if (omx > mx || (omx == mx && oam < am)) { mx = omx; am = oam; }
        }
        float se = 0.f;
        for (int j = lane; j < n; j += 32) se += __expf(lr[o + j] - mx);
        se = warp_sum(se);
        const float lse = mx + __logf(se);
        const bool t_ok = (t >= 0 && t < n);
        const float lt = t_ok ? lr[o + t] : 0.f;
        if (lane == 0) {
          if (mk != 0.f && t_ok) acc_loss[s] += (lse - lt) * mk;
          if (t_ok && am == t) acc_cor[s] += mk;
          if (argmax_out) argmax_out[row * meta.nseg + s] = am;
        }
        if (dlogits) {
          const float coef = (mk != 0.f) ? mk * meta.w[s] / (sum_w * den[s]) * grad_scale : 0.f;
          const float inv = 1.f / se;
          for (int j = lane; j < n; j += 32) {
            float g = 0.f;
            if (coef != 0.f) g = (__expf(lr[o + j] - mx) * inv - ((j == t) ? 1.f : 0.f)) * coef;
            dlogits[row * V + o + j] = from_f<T>(g);
          }
        }
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      if (s < meta.nseg) {
        if (acc_loss[s] != 0.f) atomicAdd(loss_num + s, acc_loss[s]);
        if (acc_cor[s] != 0.f) atomicAdd(correct + s, acc_cor[s]);
      }
    }
  }
}

// den[s] = sum_m mask[m, s]
__global__ void __launch_bounds__(256) mask_sums_kernel(const float* __restrict__ mask, float* __restrict__ den,
                                                        long long M, int nseg) {
  pdl_entry();
  float acc = 0.f;
  const long long total = M * nseg;
  // thread handles a fixed segment: stride by a multiple of nseg
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long stride_al = stride - (stride % nseg);
  if (tid >= stride_al) return;
  const int s = (int)(tid % nseg);
  for (long long i = tid; i < total; i += stride_al) acc += mask[i];
  atomicAdd(den + s, acc);
}

// ------------------------------------------------------------------ optimizer
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  pdl_entry();
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) acc += g[i] * g[i];
  acc = warp_sum(acc);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    atomicAdd(out, t);
  }
}

// HF transformers AdamW semantics (reference pretrain.py:76 `AdamW(lr, weight_decay=0.01)`, transformers
// 4.29.2 optimization.py): eps added to sqrt(v) before bias correction, bias correction folded into the
// step size, decoupled decay applied after the update with plain lr.  Gradient clipping
// (pretrain.py:195, torch clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6))) is folded in:
// *gnorm_sq holds the squared global norm; max_norm <= 0 disables clipping.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                    const float* __restrict__ g, bf16* __restrict__ p_bf16, long long n,
                                                    float lr, float beta1, float beta2, float eps, float wd,
                                                    float step_size, const float* __restrict__ gnorm_sq, float max_norm,
                                                    float grad_scale, float bf16_scale) {
  pdl_entry();
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float norm = sqrtf(*gnorm_sq) * grad_scale;
    coef *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  auto upd = [&](float& pi, float& mi, float& vi, float gi) {
    gi *= coef;
    mi = beta1 * mi + (1.f - beta1) * gi;
    vi = beta2 * vi + (1.f - beta2) * gi * gi;
    pi = pi - step_size * mi / (sqrtf(vi) + eps);
    pi -= lr * wd * pi;
  };
  // 16-byte accesses, two independent packs per thread and iteration (30 bytes move per element: the kernel is pure HBM
  // streaming and needs the loads of a whole iteration in flight at once); callers pass 16-byte aligned pointers
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < n4;
    float4 P0 = reinterpret_cast<const float4*>(p)[i], M0 = reinterpret_cast<const float4*>(m)[i];
    float4 V0 = reinterpret_cast<const float4*>(v)[i], G0 = reinterpret_cast<const float4*>(g)[i];
    float4 P1, M1, V1, G1;
    if (two) {
      P1 = reinterpret_cast<const float4*>(p)[i2]; M1 = reinterpret_cast<const float4*>(m)[i2];
      V1 = reinterpret_cast<const float4*>(v)[i2]; G1 = reinterpret_cast<const float4*>(g)[i2];
    }
    upd(P0.x, M0.x, V0.x, G0.x); upd(P0.y, M0.y, V0.y, G0.y); upd(P0.z, M0.z, V0.z, G0.z); upd(P0.w, M0.w, V0.w, G0.w);
    reinterpret_cast<float4*>(p)[i] = P0; reinterpret_cast<float4*>(m)[i] = M0; reinterpret_cast<float4*>(v)[i] = V0;
    if (p_bf16) {
      uint2 o;
      *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(P0.x * bf16_scale, P0.y * bf16_scale);
      *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(P0.z * bf16_scale, P0.w * bf16_scale);
      reinterpret_cast<uint2*>(p_bf16)[i] = o;
    }
    if (two) {
      upd(P1.x, M1.x, V1.x, G1.x); upd(P1.y, M1.y, V1.y, G1.y); upd(P1.z, M1.z, V1.z, G1.z); upd(P1.w, M1.w, V1.w, G1.w);
      reinterpret_cast<float4*>(p)[i2] = P1; reinterpret_cast<float4*>(m)[i2] = M1; reinterpret_cast<float4*>(v)[i2] = V1;
      if (p_bf16) {
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(P1.x * bf16_scale, P1.y * bf16_scale);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(P1.z * bf16_scale, P1.w * bf16_scale);
        reinterpret_cast<uint2*>(p_bf16)[i2] = o;
      }
    }
  }
  // tail (n not a multiple of 4)
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, mi, vi, g[i]);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (p_bf16) p_bf16[i] = __float2bfloat16(pi * bf16_scale);
  }
}

// y[m, :] = x[m, :] + table[m % S, :]   (external decoder input embeddings + learned positions)
template <typename T>
__global__ void __launch_bounds__(256) add_rows_mod_kernel(const T* __restrict__ x, const T* __restrict__ table, T* __restrict__ y,
                                                           long long M, int d, int S) {
  pdl_entry();
  constexpr int N = Pack<T>::N;
  const long long packs = M * (d / N);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < packs; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / (d / N);
    const int c = (int)(i % (d / N)) * N;
    float a[N], b[N];
    load_pack(x + m * d + c, a);
    load_pack(table + (m % S) * d + c, b);
#pragma unroll
    for (int j = 0; j < N; ++j) a[j] += b[j];
    store_pack(y + m * d + c, a);
  }
}

template <typename TO>
__global__ void __launch_bounds__(256) cast_scale_kernel(const float* __restrict__ src, TO* __restrict__ dst, long long n,
                                                         float scale) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f<TO>(src[i] * scale);
}
template <typename TI>
__global__ void __launch_bounds__(256) to_f32_kernel(const TI* __restrict__ src, float* __restrict__ dst, long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = to_f(src[i]);
}

}  // namespace

#define PB_STREAM(s) reinterpret_cast<cudaStream_t>(s)

static int fill_embed_meta(EmbedMeta& meta, const int* n_tokens_host) {
  int off = 0;
  for (int i = 0; i < 8; ++i) { meta.row_off[i] = off; meta.n_tok[i] = n_tokens_host[i]; off += n_tokens_host[i]; }
  return off;
}

extern "C" int pb_octuple_embed_fwd(const void* ids, int ids_int64, const void* table, void* out, long long M,
                                    const int* n_tokens_host, int dtype, int* err_flag, void* stream) {
  EmbedMeta meta;
  fill_embed_meta(meta, n_tokens_host);
  const int grid = grid_for(M, 1, 16);
  if (dtype == PB_DTYPE_BF16) {
    if (ids_int64) PB_LAUNCH((octuple_embed_fwd_kernel<bf16, long long>), grid, 256, 0, PB_STREAM(stream), (const long long*)ids, (const bf16*)table, (bf16*)out, M, meta, err_flag);
    else PB_LAUNCH((octuple_embed_fwd_kernel<bf16, int>), grid, 256, 0, PB_STREAM(stream), (const int*)ids, (const bf16*)table, (bf16*)out, M, meta, err_flag);
  } else {
    if (ids_int64) PB_LAUNCH((octuple_embed_fwd_kernel<float, long long>), grid, 256, 0, PB_STREAM(stream), (const long long*)ids, (const float*)table, (float*)out, M, meta, err_flag);
    else PB_LAUNCH((octuple_embed_fwd_kernel<float, int>), grid, 256, 0, PB_STREAM(stream), (const int*)ids, (const float*)table, (float*)out, M, meta, err_flag);
  }
  return pb_check_launch("octuple_embed_fwd");
}

extern "C" int pb_octuple_embed_bwd(const void* ids, int ids_int64, const void* dx, float* dtable, long long M,
                                    const int* n_tokens_host, float scale, int dtype, void* stream) {
  EmbedMeta meta;
  fill_embed_meta(meta, n_tokens_host);
  int total_rows = 0;
  for (int i = 0; i < 8; ++i) total_rows += n_tokens_host[i];
  const int smem = total_rows * EB_COLS * (int)sizeof(float);
  if (smem > 200 * 1024) return pb_set_error("octuple_embed_bwd: vocabulary too large for the shared-memory accumulator");
  long long chunks = (2LL * pb_num_sms() + 7) / 8;                 // 8 column slices x chunks ~ two waves of CTAs at most
  if (smem > 100 * 1024) chunks = pb_num_sms() / 8;                // one CTA per SM fits: a single wave
  if (chunks > (M + 7) / 8) chunks = (M + 7) / 8;
  if (chunks < 1) chunks = 1;
  const long long per = (M + chunks - 1) / chunks;
  const dim3 grid(256 / EB_COLS, (unsigned)chunks);
#define PB_EB_LAUNCH(TT, II)                                                                                          \
  do {                                                                                                               \
    static bool attr_done = false;                                                                                   \
    if (!attr_done) {                                                                                                \
      cudaFuncSetAttribute(octuple_embed_bwd_kernel<TT, II>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      attr_done = true;                                                                                              \
    }                                                                                                                \
    PB_LAUNCH((octuple_embed_bwd_kernel<TT, II>), grid, 256, smem, PB_STREAM(stream), (const II*)ids, (const TT*)dx, dtable, M, \
              meta, scale, total_rows, per);                                                                         \
  } while (0)
  if (dtype == PB_DTYPE_BF16) {
    if (ids_int64) PB_EB_LAUNCH(bf16, long long); else PB_EB_LAUNCH(bf16, int);
  } else {
    if (ids_int64) PB_EB_LAUNCH(float, long long); else PB_EB_LAUNCH(float, int);
  }
#undef PB_EB_LAUNCH
  return pb_check_launch("octuple_embed_bwd");
}

static int ln_packs(int d, int dtype) {
  const int n = dtype == PB_DTYPE_BF16 ? 8 : 4;
  if (d % n != 0 || d > 8 * 32 * n) {
    pb_set_error("layernorm: unsupported width (needs d % pack == 0 and d <= 1024 fp32 / 2048 bf16)");
    return -1;
  }
  const int packs = (d / n + 31) / 32;
  return packs <= 1 ? 1 : (packs <= 2 ? 2 : (packs <= 4 ? 4 : 8));
}

template <typename T, int MAXP>
static void ln_fwd_launch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                          long long M, int d, float eps, pbdrop::Site drop, cudaStream_t st) {
  const int grid = grid_for(M, 4, 16);
  PB_LAUNCH((layernorm_fwd_kernel<T, MAXP>), grid, 128, 0, st, (const T*)x, gamma, beta, (T*)y, mean, rstd, M, d, eps, drop);
}
template <typename T, int MAXP>
static void ln_bwd_launch(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                          void* dx_drop, float* dgamma, float* dbeta, float* dbias, long long M, int d, pbdrop::Site din,
                          pbdrop::Site dout, cudaStream_t st) {
  // one 16-warp block per SM: the per-column partial sums are reduced through shared memory first, so only
  // #SM atomics per column reach L2 (the 12 KB gradient row is a contention hot spot otherwise)
  const int grid = grid_for(M, LNB_WARPS * 2, 1);
  const int smem = LNB_WARPS * d * (int)sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_bwd_kernel<T, MAXP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  PB_LAUNCH((layernorm_bwd_kernel<T, MAXP>), grid, LNB_WARPS * 32, smem, st, (const T*)dy, (const T*)x, gamma, mean, rstd, (T*)dx,
                                                                 (T*)dx_drop, dgamma, dbeta, dbias, M, d, din, dout);
}

#define LN_DISPATCH(FN, ...)                                                         \
  if (dtype == PB_DTYPE_BF16) {                                                      \
    switch (mp) { case 1: FN<bf16, 1>(__VA_ARGS__); break; case 2: FN<bf16, 2>(__VA_ARGS__); break; \
                  case 4: FN<bf16, 4>(__VA_ARGS__); break; default: FN<bf16, 8>(__VA_ARGS__); }      \
  } else {                                                                           \
    switch (mp) { case 1: FN<float, 1>(__VA_ARGS__); break; case 2: FN<float, 2>(__VA_ARGS__); break; \
                  case 4: FN<float, 4>(__VA_ARGS__); break; default: FN<float, 8>(__VA_ARGS__); }     \
  }

static pbdrop::Site to_site(const pb_drop_site* s) {
  pbdrop::Site r;
  r.seed = s ? s->seed : nullptr; r.op = s ? s->op : 0; r.thresh = s ? s->thresh : 0; r.scale = s ? s->scale : 1.f;
  return r;
}

extern "C" int pb_layernorm_fwd_drop(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                     long long M, int d, float eps, const pb_drop_site* out_site, int dtype, void* stream) {
  const int mp = ln_packs(d, dtype);
  if (mp < 0) return -1;
  const pbdrop::Site ds = to_site(out_site);
  LN_DISPATCH(ln_fwd_launch, x, gamma, beta, y, mean, rstd, M, d, eps, ds, PB_STREAM(stream));
  return pb_check_launch("layernorm_fwd");
}
extern "C" int pb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                long long M, int d, float eps, int dtype, void* stream) {
  return pb_layernorm_fwd_drop(x, gamma, beta, y, mean, rstd, M, d, eps, nullptr, dtype, stream);
}

extern "C" int pb_layernorm_bwd_drop(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                     void* dx, void* dx_drop, float* dgamma, float* dbeta, float* dbias, long long M, int d,
                                     const pb_drop_site* in_site, const pb_drop_site* out_site, int dtype, void* stream) {
  const int mp = ln_packs(d, dtype);
  if (mp < 0) return -1;
  const pbdrop::Site di = to_site(in_site), dso = to_site(out_site);
  if (dso.seed && dx_drop == nullptr) return pb_set_error("layernorm_bwd: out dropout site needs dx_drop");
  LN_DISPATCH(ln_bwd_launch, dy, x, gamma, mean, rstd, dx, dx_drop, dgamma, dbeta, dbias, M, d, di, dso, PB_STREAM(stream));
  return pb_check_launch("layernorm_bwd");
}
extern "C" int pb_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                void* dx, float* dgamma, float* dbeta, float* dbias, long long M, int d, int dtype,
                                void* stream) {
  return pb_layernorm_bwd_drop(dy, x, gamma, mean, rstd, dx, nullptr, dgamma, dbeta, dbias, M, d, nullptr, nullptr, dtype, stream);
}

__global__ void dropout_mask_kernel(const unsigned long long* seed, uint32_t op, uint32_t thresh, unsigned char* mask, long long n) {
  pdl_entry();
  const uint32_t key = pbdrop::site_key(*seed, op);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    mask[i] = pbdrop::keep(key, (unsigned long long)i, thresh) ? 1 : 0;
}
extern "C" int pb_dropout_mask(const unsigned long long* seed, unsigned int op, unsigned int thresh, unsigned char* mask,
                               long long n, void* stream) {
  PB_LAUNCH((dropout_mask_kernel), grid_for(n, 256 * 8, 8), 256, 0, PB_STREAM(stream), seed, op, thresh, mask, n);
  return pb_check_launch("dropout_mask");
}
__global__ void add_u64_kernel(unsigned long long* p, unsigned long long inc) {
  pdl_entry(); *p += inc; }
extern "C" int pb_add_u64(unsigned long long* ptr, unsigned long long inc, void* stream) {
  PB_LAUNCH((add_u64_kernel), 1, 1, 0, PB_STREAM(stream), ptr, inc);
  return pb_check_launch("add_u64");
}

extern "C" int pb_softmax_fwd(const float* scores, void* probs, const uint8_t* key_keep, int B, int H, int Sq, int Sk,
                              int causal, int dtype, void* stream) {
  if (Sk > SM_MAX * 32) return pb_set_error("softmax: Sk > 1024 not supported");
  const int grid = grid_for((long long)B * H * Sq, 4, 16);
  if (dtype == PB_DTYPE_BF16)
    PB_LAUNCH((softmax_fwd_kernel<bf16>), grid, 128, 0, PB_STREAM(stream), scores, (bf16*)probs, key_keep, B, H, Sq, Sk, causal);
  else
    PB_LAUNCH((softmax_fwd_kernel<float>), grid, 128, 0, PB_STREAM(stream), scores, (float*)probs, key_keep, B, H, Sq, Sk, causal);
  return pb_check_launch("softmax_fwd");
}

extern "C" int pb_softmax_bwd(const void* probs, const float* dprobs, void* dscores, const uint8_t* key_keep, int B, int H,
                              int Sq, int Sk, int causal, int dtype, void* stream) {
  if (Sk > SM_MAX * 32) return pb_set_error("softmax: Sk > 1024 not supported");
  const int grid = grid_for((long long)B * H * Sq, 4, 16);
  if (dtype == PB_DTYPE_BF16)
    PB_LAUNCH((softmax_bwd_kernel<bf16>), grid, 128, 0, PB_STREAM(stream), (const bf16*)probs, dprobs, (bf16*)dscores, key_keep, B, H, Sq, Sk, causal);
  else
    PB_LAUNCH((softmax_bwd_kernel<float>), grid, 128, 0, PB_STREAM(stream), (const float*)probs, dprobs, (float*)dscores, key_keep, B, H, Sq, Sk, causal);
  return pb_check_launch("softmax_bwd");
}

extern "C" int pb_colsum(const void* x, float* out, long long M, int N, long long ld, int dtype, void* stream) {
  const int pn = dtype == PB_DTYPE_BF16 ? 8 : 4;
  if (N % pn != 0 || ld % pn != 0) return pb_set_error("colsum: N and ld must be multiples of the pack width");
  const int gx = (N + 32 * pn - 1) / (32 * pn);
  int rows_per_block = 128;
  // keep roughly >= 4 waves of blocks without exploding the atomic count
  while (rows_per_block > 16 && (long long)gx * ((M + rows_per_block - 1) / rows_per_block) < 4LL * pb_num_sms()) rows_per_block >>= 1;
  const long long gy = (M + rows_per_block - 1) / rows_per_block;
  if (gy > 65535) return pb_set_error("colsum: too many row blocks");
  dim3 grid(gx, (unsigned)gy);
  if (dtype == PB_DTYPE_BF16) PB_LAUNCH((colsum_kernel<bf16>), grid, 256, 0, PB_STREAM(stream), (const bf16*)x, out, M, N, ld, rows_per_block);
  else PB_LAUNCH((colsum_kernel<float>), grid, 256, 0, PB_STREAM(stream), (const float*)x, out, M, N, ld, rows_per_block);
  return pb_check_launch("colsum");
}

extern "C" int pb_heads_ce(const float* logits, const int* targets, const float* mask, const float* den, float* loss_num,
                           float* correct, void* dlogits, int* argmax_out, long long M, int nseg,
                           const int* seg_sizes_host, const float* weights_host, float grad_scale, int dtype,
                           void* stream) {
  if (nseg < 1 || nseg > 16) return pb_set_error("heads_ce: nseg must be in [1,16]");
  SegMeta meta;
  meta.nseg = nseg;
  int off = 0;
  float sw = 0.f;
  for (int s = 0; s < nseg; ++s) { meta.off[s] = off; off += seg_sizes_host[s]; meta.w[s] = weights_host[s]; sw += weights_host[s]; }
  meta.off[nseg] = off;
  const int grid = grid_for(M, 4 * 4, 8);
  if (dtype == PB_DTYPE_BF16)
    PB_LAUNCH((heads_ce_kernel<bf16>), grid, 128, 0, PB_STREAM(stream), logits, targets, mask, den, loss_num, correct, (bf16*)dlogits, argmax_out, M, meta, sw, grad_scale);
  else
    PB_LAUNCH((heads_ce_kernel<float>), grid, 128, 0, PB_STREAM(stream), logits, targets, mask, den, loss_num, correct, (float*)dlogits, argmax_out, M, meta, sw, grad_scale);
  return pb_check_launch("heads_ce");
}

extern "C" int pb_mask_sums(const float* mask, float* den, long long M, int nseg, void* stream) {
  PB_LAUNCH((mask_sums_kernel), grid_for(M * nseg, 256 * 8, 4), 256, 0, PB_STREAM(stream), mask, den, M, nseg);
  return pb_check_launch("mask_sums");
}

extern "C" int pb_sumsq(const float* g, long long n, float* out, void* stream) {
  if ((reinterpret_cast<uintptr_t>(g) & 15) != 0) return pb_set_error("sumsq: pointer not 16-byte aligned");
  PB_LAUNCH((sumsq_kernel), grid_for(n, 256 * 16, 8), 256, 0, PB_STREAM(stream), g, n, out);
  return pb_check_launch("sumsq");
}

extern "C" int pb_adamw(float* p, float* m, float* v, const float* g, void* p_bf16, long long n, float lr, float beta1,
                        float beta2, float eps, float wd, int step, const float* gnorm_sq, float max_norm,
                        float grad_scale, float bf16_scale, void* stream) {
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(g)) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(p_bf16) & 7) != 0)
    return pb_set_error("adamw: pointers must be 16-byte aligned (bf16 copy: 8-byte)");
  PB_LAUNCH((adamw_kernel), grid_for(n, 256 * 8, 8), 256, 0, PB_STREAM(stream), p, m, v, g, (bf16*)p_bf16, n, lr, beta1, beta2, eps, wd, step_size, gnorm_sq, max_norm, grad_scale, bf16_scale);
  return pb_check_launch("adamw");
}

extern "C" int pb_add_rows_mod(const void* x, const void* table, void* y, long long M, int d, int S, int dtype, void* stream) {
  const int pn = dtype == PB_DTYPE_BF16 ? 8 : 4;
  if (d % pn != 0) return pb_set_error("add_rows_mod: d must be a multiple of the pack width");
  const int grid = grid_for(M * (d / pn), 256 * 4, 8);
  if (dtype == PB_DTYPE_BF16) PB_LAUNCH((add_rows_mod_kernel<bf16>), grid, 256, 0, PB_STREAM(stream), (const bf16*)x, (const bf16*)table, (bf16*)y, M, d, S);
  else PB_LAUNCH((add_rows_mod_kernel<float>), grid, 256, 0, PB_STREAM(stream), (const float*)x, (const float*)table, (float*)y, M, d, S);
  return pb_check_launch("add_rows_mod");
}

extern "C" int pb_cast_from_f32(const float* src, void* dst, long long n, float scale, int dtype, void* stream) {
  const int grid = grid_for(n, 256 * 8, 8);
  if (dtype == PB_DTYPE_BF16) PB_LAUNCH((cast_scale_kernel<bf16>), grid, 256, 0, PB_STREAM(stream), src, (bf16*)dst, n, scale);
  else PB_LAUNCH((cast_scale_kernel<float>), grid, 256, 0, PB_STREAM(stream), src, (float*)dst, n, scale);
  return pb_check_launch("cast_from_f32");
}

extern "C" int pb_cast_to_f32(const void* src, float* dst, long long n, int dtype, void* stream) {
  const int grid = grid_for(n, 256 * 8, 8);
  if (dtype == PB_DTYPE_BF16) PB_LAUNCH((to_f32_kernel<bf16>), grid, 256, 0, PB_STREAM(stream), (const bf16*)src, dst, n);
  else PB_LAUNCH((to_f32_kernel<float>), grid, 256, 0, PB_STREAM(stream), (const float*)src, dst, n);
  return pb_check_launch("cast_to_f32");
}

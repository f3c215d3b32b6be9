// Fused Octuple front end (north-star subsystem 1; reference PianoBart.py:9-16,60-71 + HF BartEncoder/BartDecoder prologue
// modeling_bart.py:521-526,649-655):
//
//   y0[m] = in_linear(cat_a(E_a[id_a[m]] * 16)) + pos[m % S]        h0[m] = dropout(LayerNorm(y0[m]))
//
// The concatenated embedding X [M, 2048] never exists.  in_linear is linear in the concatenation, so
//   in_linear(cat_a e_a) = sum_a  e_a W_a^T + b,      W_a = W[:, 256 a : 256 a + 256],
// and because the Octuple vocabulary is tiny (1280 rows over the 8 attributes) the products are tabulated once per step:
//   T[off_a + r] = 16 E_a[r] W_a^T                      ([1280, d] in the activation dtype: 2.6 MB bf16, L2-resident)
// (one tcgen05 GEMM over the block-diagonal copy of the tables, octuple_blockdiag_kernel below).  One kernel then gathers 8 rows of T per token, adds bias and the
// position row, writes the pre-LayerNorm sum (LayerNorm's backward input) and the normalised, dropped-out output: one warp
// per token, the row lives in registers, 16-byte accesses.  HBM traffic per token and stream: 8 ids in, 2 x d x 2 B out
// (the 8 x d x 2 B of table rows come from L2), instead of the 2 x 4 KB round trip of X plus a [M,2048] x [2048,d] GEMM.
//
// Backward (engine.py): with G = sum_m onehot(m)^T dy0[m] ([1280, d], one tcgen05 GEMM against the one-hot matrix built by
// octuple_onehot_kernel),  dE_a = 16 G_a W_a  and  dW_a = G_a^T (16 E_a)  are two more GEMMs (block-diagonal form) - the [M,2048] gradient of
// X, the [M,2048] x [d,2048] weight-gradient product and the scatter-add of round 1 (4.5 % of HBM peak) disappear.
#include "pb_internal.h"
#include "dropout.cuh"
#include <cuda_bf16.h>
#include <stdint.h>

namespace {

typedef __nv_bfloat16 bf16;

template <typename T> struct Pk;
template <> struct Pk<float> { static constexpr int N = 4; typedef float4 raw; };
template <> struct Pk<bf16> { static constexpr int N = 8; typedef uint4 raw; };

__device__ __forceinline__ void unpack(const float4& r, float (&f)[4]) { f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w; }
__device__ __forceinline__ void unpack(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
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
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct FrontMeta { int row_off[8]; int n_tok[8]; };

template <typename T, typename I, int MAXP>
__global__ void __launch_bounds__(128) octuple_front_fwd_kernel(const I* __restrict__ ids, const T* __restrict__ table,
                                                                const float* __restrict__ bias, const T* __restrict__ pos, int S,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                T* __restrict__ y0, T* __restrict__ h0, float* __restrict__ mean_out,
                                                                float* __restrict__ rstd_out, long long M, int d, float eps,
                                                                FrontMeta meta, pbdrop::Site drop, int* __restrict__ err) {
  pdl_entry();
  constexpr int N = Pk<T>::N;
  typedef typename Pk<T>::raw Raw;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const uint32_t dkey = drop.seed ? pbdrop::site_key(*drop.seed, drop.op) : 0u;
  __shared__ __align__(16) float s_ga[MAXP * 32 * N], s_be[MAXP * 32 * N], s_bi[MAXP * 32 * N];
  for (int c = threadIdx.x; c < MAXP * 32 * N; c += blockDim.x) {
    s_ga[c] = c < d ? __ldg(gamma + c) : 0.f;
    s_be[c] = c < d ? __ldg(beta + c) : 0.f;
    s_bi[c] = c < d ? __ldg(bias + c) : 0.f;
  }
  __syncthreads();
  for (long long row = warp_global; row < M; row += nwarps) {
    // the 8 table rows of this token (lanes 0-7 read and validate one id each)
    int trow = 0;
    if (lane < 8) {
      long long id = (long long)ids[row * 8 + lane];
      if (id < 0 || id >= meta.n_tok[lane]) { if (err) atomicExch(err, 1); id = 0; }
      trow = meta.row_off[lane] + (int)id;
    }
    int tr[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) tr[a] = __shfl_sync(0xffffffffu, trow, a);
    const long long prow = (long long)(row % S) * d;
    float v[MAXP][N];
    float s = 0.f;
    T* yr = y0 + row * d;
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) {
        Raw tv[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) tv[a] = *reinterpret_cast<const Raw*>(table + (long long)tr[a] * d + c);
        const Raw pv = *reinterpret_cast<const Raw*>(pos + prow + c);
        float f[N];
#pragma unroll
        for (int j = 0; j < N; ++j) v[k][j] = s_bi[c + j];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          unpack(tv[a], f);
#pragma unroll
          for (int j = 0; j < N; ++j) v[k][j] += f[j];
        }
        unpack(pv, f);
#pragma unroll
        for (int j = 0; j < N; ++j) v[k][j] += f[j];
        store_pack(yr + c, v[k]);
        // LayerNorm sees exactly what its backward will re-read (the value rounded to the activation dtype)
        if constexpr (sizeof(T) == 2) {
#pragma unroll
          for (int j = 0; j < N; ++j) v[k][j] = __bfloat162float(__float2bfloat16(v[k][j]));
        }
#pragma unroll
        for (int j = 0; j < N; ++j) s += v[k][j];
      }
    }
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
    T* hr = h0 + row * d;
#pragma unroll
    for (int k = 0; k < MAXP; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < d) {
        float o[N];
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = (v[k][j] - mean) * rstd * s_ga[c + j] + s_be[c + j];
        if (drop.seed) {
          const uint32_t bits = pbdrop::keep_bits<N>(dkey, (unsigned long long)row * d + c, drop.thresh);
#pragma unroll
          for (int j = 0; j < N; ++j) o[j] = ((bits >> j) & 1u) ? o[j] * drop.scale : 0.f;
        }
        store_pack(hr + c, o);
      }
    }
  }
}

// out[m, off_a + ids[m, a]] = 1 for the 8 attributes, 0 elsewhere ([M, V] in the activation dtype; V = total rows, V % N == 0)
template <typename T, typename I>
__global__ void __launch_bounds__(256) octuple_onehot_kernel(const I* __restrict__ ids, T* __restrict__ out, long long M, int V,
                                                             FrontMeta meta) {
  pdl_entry();
  constexpr int N = Pk<T>::N;
  const int packs = V / N;
  const long long total = M * packs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / packs;
    const int c0 = (int)(i - m * packs) * N;
    float f[N];
#pragma unroll
    for (int j = 0; j < N; ++j) f[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const long long id = (long long)ids[m * 8 + a];
      if (id < 0 || id >= meta.n_tok[a]) continue;
      const int r = meta.row_off[a] + (int)id - c0;
      if (r >= 0 && r < N) {
#pragma unroll
        for (int j = 0; j < N; ++j) if (j == r) f[j] = 1.f;
      }
    }
    store_pack(out + m * V + c0, f);
  }
}

// Block-diagonal form of the eight embedding tables: out[off_a + r, E a + c] = emb[off_a + r, c]   ([V, 8 E], zero elsewhere -
// the off-diagonal blocks are cleared once by the caller and never written).  With it the eight table products of the front
// end become ONE GEMM  T = Ebd W_in^T  (the zero blocks add exact zeros to the fp32 accumulators, so T is bit-identical to
// the per-attribute products), and the eight weight-gradient products ONE:  dW_in += G^T Ebd.
template <typename T>
__global__ void __launch_bounds__(256) octuple_blockdiag_kernel(const T* __restrict__ emb, T* __restrict__ out, int V, int E,
                                                                FrontMeta meta) {
  pdl_entry();
  constexpr int N = Pk<T>::N;
  typedef typename Pk<T>::raw Raw;
  const int packs = E / N;
  const int total = V * packs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / packs, c = (i - r * packs) * N;
    int a = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) a += (r >= meta.row_off[k]) ? 1 : 0;
    *reinterpret_cast<Raw*>(out + (long long)r * (8 * E) + a * E + c) = *reinterpret_cast<const Raw*>(emb + (long long)r * E + c);
  }
}

// Gradient of the tables from the full product dEbd = G W_in ([V, 8 E] fp32): g_emb[off_a + r, c] += alpha dEbd[off_a + r, E a + c]
__global__ void __launch_bounds__(256) octuple_blockdiag_grad_kernel(const float* __restrict__ dfull, float* __restrict__ g_emb,
                                                                     int V, int E, float alpha, FrontMeta meta) {
  pdl_entry();
  const int packs = E / 4;
  const int total = V * packs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / packs, c = (i - r * packs) * 4;
    int a = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) a += (r >= meta.row_off[k]) ? 1 : 0;
    const float4 v = *reinterpret_cast<const float4*>(dfull + (long long)r * (8 * E) + a * E + c);
    float4* g = reinterpret_cast<float4*>(g_emb + (long long)r * E + c);
    float4 o = *g;
    o.x += alpha * v.x; o.y += alpha * v.y; o.z += alpha * v.z; o.w += alpha * v.w;
    *g = o;
  }
}

FrontMeta make_meta(const int* n_tokens_host, int& total) {
  FrontMeta m;
  int off = 0;
  for (int i = 0; i < 8; ++i) { m.row_off[i] = off; m.n_tok[i] = n_tokens_host[i]; off += n_tokens_host[i]; }
  total = off;
  return m;
}

inline int grid_rows(long long rows, int per_block, int max_per_sm) {
  long long b = (rows + per_block - 1) / per_block;
  const long long cap = (long long)pb_num_sms() * max_per_sm;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

template <typename T, typename I, int MAXP>
void front_launch(const void* ids, const void* table, const float* bias, const void* pos, int S, const float* gamma,
                  const float* beta, void* y0, void* h0, float* mean, float* rstd, long long M, int d, float eps,
                  const FrontMeta& meta, pbdrop::Site ds, int* err, cudaStream_t st) {
  PB_LAUNCH((octuple_front_fwd_kernel<T, I, MAXP>), grid_rows(M, 4, 12), 128, 0, st, (const I*)ids, (const T*)table, bias,
            (const T*)pos, S, gamma, beta, (T*)y0, (T*)h0, mean, rstd, M, d, eps, meta, ds, err);
}

}  // namespace

#define PB_STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int pb_octuple_front_fwd(const void* ids, int ids_int64, const void* table_proj, const float* bias,
                                    const void* pos_rows, int S, const float* gamma, const float* beta, void* y0, void* h0,
                                    float* mean, float* rstd, long long M, int d, const int* n_tokens_host, float eps,
                                    const pb_drop_site* out_site, int dtype, int* err_flag, void* stream) {
  int total;
  const FrontMeta meta = make_meta(n_tokens_host, total);
  const int n = dtype == PB_DTYPE_BF16 ? 8 : 4;
  if (d % n != 0 || S <= 0) return pb_set_error("octuple_front_fwd: d must be a multiple of the 16-byte pack");
  int mp = (d + 32 * n - 1) / (32 * n);
  if (mp > 8) return pb_set_error("octuple_front_fwd: d too large (<= 2048 bf16 / 1024 fp32)");
  mp = mp <= 1 ? 1 : (mp <= 2 ? 2 : (mp <= 4 ? 4 : 8));
  pbdrop::Site ds;
  ds.seed = out_site ? out_site->seed : nullptr; ds.op = out_site ? out_site->op : 0;
  ds.thresh = out_site ? out_site->thresh : 0; ds.scale = out_site ? out_site->scale : 1.f;
  cudaStream_t st = PB_STREAM(stream);
#define FRONT_GO(T, I)                                                                                                   \
  switch (mp) {                                                                                                          \
    case 1: front_launch<T, I, 1>(ids, table_proj, bias, pos_rows, S, gamma, beta, y0, h0, mean, rstd, M, d, eps, meta, ds, err_flag, st); break; \
    case 2: front_launch<T, I, 2>(ids, table_proj, bias, pos_rows, S, gamma, beta, y0, h0, mean, rstd, M, d, eps, meta, ds, err_flag, st); break; \
    case 4: front_launch<T, I, 4>(ids, table_proj, bias, pos_rows, S, gamma, beta, y0, h0, mean, rstd, M, d, eps, meta, ds, err_flag, st); break; \
    default: front_launch<T, I, 8>(ids, table_proj, bias, pos_rows, S, gamma, beta, y0, h0, mean, rstd, M, d, eps, meta, ds, err_flag, st); }
  if (dtype == PB_DTYPE_BF16) {
    if (ids_int64) { FRONT_GO(bf16, long long) } else { FRONT_GO(bf16, int) }
  } else {
    if (ids_int64) { FRONT_GO(float, long long) } else { FRONT_GO(float, int) }
  }
#undef FRONT_GO
  return pb_check_launch("octuple_front_fwd");
}

extern "C" int pb_octuple_onehot(const void* ids, int ids_int64, void* out, long long M, const int* n_tokens_host, int dtype,
                                 void* stream) {
  int total;
  const FrontMeta meta = make_meta(n_tokens_host, total);
  const int n = dtype == PB_DTYPE_BF16 ? 8 : 4;
  if (total % n != 0) return pb_set_error("octuple_onehot: vocabulary size must be a multiple of the 16-byte pack");
  cudaStream_t st = PB_STREAM(stream);
  const int grid = grid_rows(M * (total / n), 256, 16);
  if (dtype == PB_DTYPE_BF16) {
    if (ids_int64) { PB_LAUNCH((octuple_onehot_kernel<bf16, long long>), grid, 256, 0, st, (const long long*)ids, (bf16*)out, M, total, meta); }
    else { PB_LAUNCH((octuple_onehot_kernel<bf16, int>), grid, 256, 0, st, (const int*)ids, (bf16*)out, M, total, meta); }
  } else {
    if (ids_int64) { PB_LAUNCH((octuple_onehot_kernel<float, long long>), grid, 256, 0, st, (const long long*)ids, (float*)out, M, total, meta); }
    else { PB_LAUNCH((octuple_onehot_kernel<float, int>), grid, 256, 0, st, (const int*)ids, (float*)out, M, total, meta); }
  }
  return pb_check_launch("octuple_onehot");
}

extern "C" int pb_octuple_blockdiag(const void* emb, void* out, int emb_dim, const int* n_tokens_host, int dtype, void* stream) {
  int total;
  const FrontMeta meta = make_meta(n_tokens_host, total);
  const int n = dtype == PB_DTYPE_BF16 ? 8 : 4;
  if (emb_dim <= 0 || emb_dim % n != 0) return pb_set_error("octuple_blockdiag: emb_dim must be a multiple of the 16-byte pack");
  cudaStream_t st = PB_STREAM(stream);
  const int grid = grid_rows((long long)total * (emb_dim / n), 256, 4);
  if (dtype == PB_DTYPE_BF16) { PB_LAUNCH((octuple_blockdiag_kernel<bf16>), grid, 256, 0, st, (const bf16*)emb, (bf16*)out, total, emb_dim, meta); }
  else { PB_LAUNCH((octuple_blockdiag_kernel<float>), grid, 256, 0, st, (const float*)emb, (float*)out, total, emb_dim, meta); }
  return pb_check_launch("octuple_blockdiag");
}

extern "C" int pb_octuple_blockdiag_grad(const float* dfull, float* g_emb, int emb_dim, const int* n_tokens_host, float alpha,
                                         void* stream) {
  int total;
  const FrontMeta meta = make_meta(n_tokens_host, total);
  if (emb_dim <= 0 || emb_dim % 4 != 0) return pb_set_error("octuple_blockdiag_grad: emb_dim must be a multiple of 4");
  const int grid = grid_rows((long long)total * (emb_dim / 4), 256, 4);
  PB_LAUNCH(octuple_blockdiag_grad_kernel, grid, 256, 0, PB_STREAM(stream), dfull, g_emb, total, emb_dim, alpha, meta);
  return pb_check_launch("octuple_blockdiag_grad");
}

extern "C" int pb_fill_zero(void* ptr, long long bytes, void* stream) {
  if (bytes <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, PB_STREAM(stream));
  if (e != cudaSuccess) return pb_set_cuda_error("pb_fill_zero", e);
  return 0;
}

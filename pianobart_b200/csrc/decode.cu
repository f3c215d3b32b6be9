// KV-cache autoregressive decode kernels (reference model.py:28-107: PianoBartLM.forward(generate=True),
// sample(), sampling(), nucleus()).  The reference re-runs the full encoder and the full 1024-position
// decoder for every generated token; here one step touches each decoder weight once (split-K tcgen05
// GEMMs with M = batch, see generate.py) plus the K/V caches, and everything that depends on the step
// index (position row, cache length, sampled token hand-over, stop flags) lives in device memory so the
// whole step is one replayable CUDA graph.
//
//   pb_decode_finalize : fp32 split-K accumulator -> (+bias, GELU, +residual, +position row, LayerNorm) -> bf16
//                        (and re-zeroes the accumulator for its next use)
//   pb_decode_attn     : one query token per (batch, head) against a K/V cache (self: append then attend to
//                        t+1 keys; cross: S_enc keys with the encoder key-padding mask)
//   pb_decode_sample   : per-attribute temperature softmax + nucleus (model.py:68-107) with host-drawn
//                        uniforms (numpy stream order preserved), greedy when p == 1 (reference quirk)
//   pb_decode_advance  : stop rule of model.py:59-65, result write, step counter increment
#include "pb_internal.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < nw; ++i) t = fmaxf(t, sh[i]);
  return t;
}

// one block per batch row
__global__ void __launch_bounds__(256) decode_finalize_kernel(float* __restrict__ acc, const float* __restrict__ bias,
                                                              const bf16* __restrict__ residual,
                                                              const bf16* __restrict__ pos_table, const int* __restrict__ t_dev,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              bf16* __restrict__ out, float* __restrict__ out_f32, int N,
                                                              int gelu, float eps) {
  extern __shared__ float row[];  // N floats
  __shared__ float red[8];
  const int b = blockIdx.x;
  float* a = acc + (long long)b * N;
  const bf16* pos = pos_table ? pos_table + (long long)(*t_dev + 2) * N : nullptr;  // BartLearnedPositionalEmbedding offset 2
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float x = a[i];
    a[i] = 0.f;
    if (bias) x += bias[i];
    if (gelu) x = gelu_erf(x);
    if (residual) x += __bfloat162float(residual[(long long)b * N + i]);
    if (pos) x += __bfloat162float(pos[i]);
    row[i] = x;
    s += x;
  }
  if (gamma) {
    const float mean = block_sum(s, red) / N;
    float q = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { const float d = row[i] - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum(q, red) / N + eps);
    for (int i = threadIdx.x; i < N; i += blockDim.x) row[i] = (row[i] - mean) * rstd * gamma[i] + beta[i];
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    if (out) out[(long long)b * N + i] = __float2bfloat16(row[i]);
    if (out_f32) out_f32[(long long)b * N + i] = row[i];
  }
}

// Split-key ("flash-decoding") single-query attention.  grid (H, B, NS), 128 threads; CTA z handles keys
// [z*128, z*128+128).  q: [B, q_ld] (head h at column h*hd).  K/V: row j of batch b at base + b*kv_bs + j*kv_ld + h*hd.
// append != 0: split 0 first copies k_new/v_new into row t = *t_dev (visible to the split that owns row t because that
// split re-reads the new row from k_new/v_new directly); keys 0..t are attended.  Otherwise n_keys keys gated by
// key_keep [B, n_keys].  Each CTA writes (max, sum, unnormalised out[hd]) to the workspace; the last CTA of a (b, h)
// group (atomic ticket) combines the partials and writes the bf16 output.
constexpr int DK = 128;  // keys per CTA
__global__ void __launch_bounds__(128) decode_attn_kernel(const bf16* __restrict__ q, int q_ld, const bf16* __restrict__ k_new,
                                                          const bf16* __restrict__ v_new, bf16* __restrict__ kc,
                                                          bf16* __restrict__ vc, long long kv_bs, int kv_ld,
                                                          const uint8_t* __restrict__ key_keep, int n_keys,
                                                          const int* __restrict__ t_dev, int append, bf16* __restrict__ out,
                                                          int out_ld, int hd, float scale, int max_keys,
                                                          float* __restrict__ ws, int* __restrict__ tickets) {
  __shared__ float qs[128];
  __shared__ float pr[DK];
  __shared__ float red[8];
  __shared__ int is_last;
  const int h = blockIdx.x, b = blockIdx.y, z = blockIdx.z, NS = gridDim.z, H = gridDim.x;
  const int tid = threadIdx.x;
  bf16* K = kc + (long long)b * kv_bs + h * hd;
  bf16* V = vc + (long long)b * kv_bs + h * hd;
  int n = n_keys, t = -1;
  if (append) {
    t = *t_dev;
    if (t >= max_keys) return;
    n = t + 1;
    if (z == 0)
      for (int i = tid; i < hd; i += blockDim.x) {
        K[(long long)t * kv_ld + i] = k_new[(long long)b * q_ld + h * hd + i];
        V[(long long)t * kv_ld + i] = v_new[(long long)b * q_ld + h * hd + i];
      }
  }
  for (int i = tid; i < hd; i += blockDim.x) qs[i] = __bfloat162float(q[(long long)b * q_ld + h * hd + i]) * scale;
  __syncthreads();
  const int j0 = z * DK;
  const int j = j0 + tid;
  const uint8_t* keep = (!append && key_keep) ? key_keep + (long long)b * n_keys : nullptr;
  // one key per thread: 2*hd bytes of the key row in 16-byte loads (row t comes from k_new: it may not be visible yet)
  float sc = -INFINITY;
  if (j < n && (!keep || keep[j])) {
    const bf16* krow = (j == t) ? (k_new + (long long)b * q_ld + h * hd) : (K + (long long)j * kv_ld);
    float d = 0.f;
    for (int i = 0; i < hd; i += 8) {
      const uint4 kv = *reinterpret_cast<const uint4*>(krow + i);
      const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kv);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = __bfloat1622float2(k2[u]);
        d += qs[i + 2 * u] * f.x + qs[i + 2 * u + 1] * f.y;
      }
    }
    sc = d;
  }
  const float mx = block_max(sc, red);
  const float e = (sc == -INFINITY) ? 0.f : __expf(sc - mx);
  pr[tid] = e;
  const float sum = block_sum(e, red);   // contains the __syncthreads that publishes pr[]
  const int cnt = min(DK, n - j0);
  float o = 0.f;
  if (tid < hd && cnt > 0) {
#pragma unroll 4
    for (int jj = 0; jj < cnt; ++jj) {
      const int key = j0 + jj;
      const bf16* vrow = (key == t) ? (v_new + (long long)b * q_ld + h * hd) : (V + (long long)key * kv_ld);
      o += pr[jj] * __bfloat162float(vrow[tid]);
    }
  }
  float* w = ws + (((long long)b * H + h) * NS + z) * (hd + 2);
  if (tid < hd) w[2 + tid] = o;
  if (tid == 0) { w[0] = mx; w[1] = sum; }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int tk = atomicAdd(tickets + b * H + h, 1);
    is_last = (tk == NS - 1);
    if (is_last) tickets[b * H + h] = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float* wb = ws + ((long long)b * H + h) * NS * (hd + 2);
  float gm = -INFINITY;
  for (int s2 = 0; s2 < NS; ++s2) gm = fmaxf(gm, wb[s2 * (hd + 2)]);
  if (tid < hd) {
    float tot = 0.f, acc = 0.f;
    for (int s2 = 0; s2 < NS; ++s2) {
      const float m2 = wb[s2 * (hd + 2)];
      const float f = (m2 == -INFINITY) ? 0.f : __expf(m2 - gm);
      tot += wb[s2 * (hd + 2) + 1] * f;
      acc += wb[s2 * (hd + 2) + 2 + tid] * f;
    }
    out[(long long)b * out_ld + h * hd + tid] = __float2bfloat16(tot > 0.f ? acc / tot : 0.f);
  }
}

struct SampleMeta { int off[9]; float temp[8]; float top_p[8]; };

// grid (8, B), 256 threads.  logits fp32 [B, 1280]; uniforms double [B, S, 8]; writes cur_tok[b, attr].
__global__ void __launch_bounds__(256) decode_sample_kernel(const float* __restrict__ logits, const double* __restrict__ uniforms,
                                                            const int* __restrict__ forced, const int* __restrict__ t_dev,
                                                            int* __restrict__ cur_tok, int* __restrict__ sampled, int S,
                                                            SampleMeta meta) {
  __shared__ float p[512];
  __shared__ float sp[512];
  __shared__ int si[512];
  __shared__ float red[8];
  const int a = blockIdx.x, b = blockIdx.y;
  const int t = *t_dev;
  if (t >= S) return;
  const int V = meta.off[8];
  const int o = meta.off[a], n = meta.off[a + 1] - meta.off[a];
  const float* lg = logits + (long long)b * V + o;
  const float invt = 1.0f / meta.temp[a];
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float x = lg[i] / meta.temp[a]; p[i] = x; mx = fmaxf(mx, x); }
  (void)invt;
  mx = block_max(mx, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float e = expf(p[i] - mx); p[i] = e; s += e; }
  s = block_sum(s, red);
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = p[i] / s;
  __syncthreads();
  __shared__ float tot_sh;
  if (threadIdx.x == 0) {
    float tot = 0.f;                       // model.py:85  probs /= (sum(probs) + 1e-5), float32 sequential sum
    for (int i = 0; i < n; ++i) tot += p[i];
    tot_sh = tot + 1e-5f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = p[i] / tot_sh;
  __syncthreads();
  // descending rank sort (np.argsort(probs)[::-1]: among equal values the higher index comes first)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = p[i];
    int r = 0;
    for (int k = 0; k < n; ++k) r += (p[k] > v) || (p[k] == v && k > i);
    sp[r] = v;
    si[r] = i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int last = 1;
    if (meta.top_p[a] < 1.0f) {            // p == 1: cumsum > 1 never holds -> top-1 (reference quirk)
      float c = 0.f;
      bool hit = false;
      for (int k = 0; k < n; ++k) { c += sp[k]; if (c > meta.top_p[a]) { last = k + 1; hit = true; break; } }
      if (!hit) last = 1;
    }
    float cs = 0.f;
    for (int k = 0; k < last; ++k) cs += sp[k];
    // np.random.choice(cand, size=1, p): cdf = cumsum(p) (double), cdf /= cdf[-1], searchsorted(u, 'right')
    double tot = 0.0;
    for (int k = 0; k < last; ++k) tot += (double)(sp[k] / cs);
    const double u = uniforms[((long long)b * S + t) * 8 + a];
    double c2 = 0.0;
    int pick = last - 1;
    for (int k = 0; k < last; ++k) { c2 += (double)(sp[k] / cs); if (c2 / tot > u) { pick = k; break; } }
    const int tok = si[pick];
    sampled[((long long)b * S + t) * 8 + a] = tok;
    cur_tok[b * 8 + a] = forced ? forced[((long long)b * S + t) * 8 + a] : tok;
  }
}

struct PadMeta { int pad[8]; };

// one thread per batch row: model.py:59-65
__global__ void decode_advance_kernel(const int* __restrict__ cur_tok, int* __restrict__ result, int* __restrict__ done,
                                      int* __restrict__ t_dev, int* __restrict__ n_written, int B, int S, PadMeta pm) {
  const int b = threadIdx.x;
  const int t = *t_dev;
  if (b < B && t < S && !done[b]) {
    bool stop = false;
    for (int a = 0; a < 8; ++a) stop |= cur_tok[b * 8 + a] >= pm.pad[a];
    if (stop) done[b] = 1;
    else {
      for (int a = 0; a < 8; ++a) result[((long long)b * S + t) * 8 + a] = cur_tok[b * 8 + a];
      n_written[b] = t + 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && t < S) *t_dev = t + 1;
}

}  // namespace

#define PB_STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int pb_decode_finalize(float* acc, const float* bias, const void* residual, const void* pos_table, const int* t_dev,
                                  const float* gamma, const float* beta, void* out, float* out_f32, int B, int N, int gelu,
                                  void* stream) {
  if (N * 4 > 48 * 1024) return pb_set_error("decode_finalize: row too wide");
  decode_finalize_kernel<<<B, 256, N * sizeof(float), PB_STREAM(stream)>>>(acc, bias, (const bf16*)residual, (const bf16*)pos_table,
                                                                          t_dev, gamma, beta, (bf16*)out, out_f32, N, gelu, 1e-5f);
  return pb_check_launch("decode_finalize");
}

extern "C" int pb_decode_attn(const void* q, int q_ld, const void* k_new, const void* v_new, void* k_cache, void* v_cache,
                              long long kv_batch_stride, int kv_ld, const uint8_t* key_keep, int n_keys, const int* t_dev,
                              int append, void* out, int out_ld, int B, int H, int hd, float scale, int max_keys,
                              float* workspace, int* tickets, void* stream) {
  if (hd > 128 || (hd % 8) != 0) return pb_set_error("decode_attn: head_dim must be a multiple of 8 and <= 128");
  const int NS = (max_keys + DK - 1) / DK;
  dim3 grid(H, B, NS);
  decode_attn_kernel<<<grid, 128, 0, PB_STREAM(stream)>>>(
      (const bf16*)q, q_ld, (const bf16*)k_new, (const bf16*)v_new, (bf16*)k_cache, (bf16*)v_cache, kv_batch_stride, kv_ld, key_keep,
      n_keys, t_dev, append, (bf16*)out, out_ld, hd, scale, max_keys, workspace, tickets);
  return pb_check_launch("decode_attn");
}

extern "C" int pb_decode_sample(const float* logits, const double* uniforms, const int* forced, const int* t_dev, int* cur_tok,
                                int* sampled, int B, int S, const int* seg_sizes_host, const float* temp_host,
                                const float* top_p_host, void* stream) {
  SampleMeta m;
  int off = 0;
  for (int i = 0; i < 8; ++i) {
    m.off[i] = off; off += seg_sizes_host[i]; m.temp[i] = temp_host[i]; m.top_p[i] = top_p_host[i];
    if (seg_sizes_host[i] > 512) return pb_set_error("decode_sample: segment > 512");
  }
  m.off[8] = off;
  dim3 grid(8, B);
  decode_sample_kernel<<<grid, 256, 0, PB_STREAM(stream)>>>(logits, uniforms, forced, t_dev, cur_tok, sampled, S, m);
  return pb_check_launch("decode_sample");
}

extern "C" int pb_decode_advance(const int* cur_tok, int* result, int* done, int* t_dev, int* n_written, int B, int S,
                                 const int* pad_host, void* stream) {
  if (B > 1024) return pb_set_error("decode_advance: batch > 1024");
  PadMeta pm;
  for (int i = 0; i < 8; ++i) pm.pad[i] = pad_host[i];
  decode_advance_kernel<<<1, ((B + 31) / 32) * 32, 0, PB_STREAM(stream)>>>(cur_tok, result, done, t_dev, n_written, B, S, pm);
  return pb_check_launch("decode_advance");
}

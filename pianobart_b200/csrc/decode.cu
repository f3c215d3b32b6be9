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
  pdl_entry();
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
                                                          float* __restrict__ ws, int* __restrict__ tickets,
                                                          float* __restrict__ out_f32) {
  pdl_entry();
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
#pragma unroll 16
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
  // P.V for this key chunk: warp w owns keys [32w, 32w+32), lane owns 4 consecutive head dims (8-byte loads,
  // 8 keys in flight), then the 4 warps are summed through shared memory
  __shared__ float po[4][128];
  {
    const int warp = tid >> 5, lane = tid & 31;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int d0 = lane * 4;
    if (d0 < hd) {
      const int jb = warp * 32, je = min(cnt, jb + 32);
#pragma unroll 32
      for (int jj = jb; jj < je; ++jj) {
        const int key = j0 + jj;
        const bf16* vrow = (key == t) ? (v_new + (long long)b * q_ld + h * hd) : (V + (long long)key * kv_ld);
        const uint2 vv = *reinterpret_cast<const uint2*>(vrow + d0);
        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vv);
        const float2 f0 = __bfloat1622float2(v2[0]), f1 = __bfloat1622float2(v2[1]);
        const float pj = pr[jj];
        a0 += pj * f0.x; a1 += pj * f0.y; a2 += pj * f1.x; a3 += pj * f1.y;
      }
      po[warp][d0] = a0; po[warp][d0 + 1] = a1; po[warp][d0 + 2] = a2; po[warp][d0 + 3] = a3;
    }
  }
  __syncthreads();
  float o = 0.f;
  if (tid < hd) o = po[0][tid] + po[1][tid] + po[2][tid] + po[3][tid];
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
    const float res = tot > 0.f ? acc / tot : 0.f;
    if (out) out[(long long)b * out_ld + h * hd + tid] = __float2bfloat16(res);
    if (out_f32) out_f32[(long long)b * out_ld + h * hd + tid] = res;
  }
}

// Skinny GEMV for small decode batches (B <= 8): every weight row is streamed from HBM exactly once by one warp
// (16-byte loads, 4 in flight per lane), all B activation vectors live in shared memory as fp32.
//   prologue (optional): x <- LayerNorm(x_raw) with (gamma, beta); CTA 0 also publishes the normalised vector
//                        (it is the residual of the next sub-layer) - this removes the separate finalize launches
//   epilogue: y = acc + bias; GELU; + residual; + position row (t+2); written as fp32 (the next kernel's raw input)
//             and/or bf16
constexpr int GV_MAXB = 8;
template <int B>
__global__ void __launch_bounds__(256) decode_gemv_kernel(const float* __restrict__ x_raw, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ x_norm_out,
                                                          const bf16* __restrict__ W, const float* __restrict__ bias,
                                                          const float* __restrict__ residual, const bf16* __restrict__ pos_table,
                                                          const int* __restrict__ t_dev, float* __restrict__ y_f32,
                                                          bf16* __restrict__ y_bf16, int N, int K, int gelu, float eps) {
  pdl_entry();
  extern __shared__ float xs[];  // B * K
  __shared__ float red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  // the weight rows do not depend on the prologue: put this warp's first row in flight now (8 x 16 B per lane
  // covers K <= 2048) so the HBM latency overlaps the activation load + LayerNorm below
  constexpr int MAXU = 8;
  uint4 wv[MAXU];
  int n = blockIdx.x * (blockDim.x >> 5) + warp;
  if (n < N) {
    const bf16* wr = W + (long long)n * K;
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int k = lane * 8 + u * 256;
      wv[u] = (k < K) ? *reinterpret_cast<const uint4*>(wr + k) : make_uint4(0, 0, 0, 0);
    }
  }
  for (int i = tid; i < B * K; i += blockDim.x) xs[i] = x_raw[i];
  __syncthreads();
  if (gamma != nullptr) {
    for (int b = 0; b < B; ++b) {
      float s = 0.f;
      for (int i = tid; i < K; i += blockDim.x) s += xs[b * K + i];
      const float mean = block_sum(s, red) / K;
      float q = 0.f;
      for (int i = tid; i < K; i += blockDim.x) { const float d = xs[b * K + i] - mean; q += d * d; }
      const float rstd = rsqrtf(block_sum(q, red) / K + eps);
      for (int i = tid; i < K; i += blockDim.x) {
        const float v = (xs[b * K + i] - mean) * rstd * gamma[i] + beta[i];
        xs[b * K + i] = v;
        if (x_norm_out != nullptr && blockIdx.x == 0) x_norm_out[b * K + i] = v;
      }
    }
    __syncthreads();
  }
  const bf16* pos = pos_table ? pos_table + (long long)(*t_dev + 2) * N : nullptr;
  for (; n < N; n += nwarps) {
    float acc[B];
#pragma unroll
    for (int b = 0; b < B; ++b) acc[b] = 0.f;
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int k = lane * 8 + u * 256;
      if (k < K) {
        const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&wv[u]);
        float wf[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(w2[j]); wf[2 * j] = f.x; wf[2 * j + 1] = f.y; }
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float4 xa = *reinterpret_cast<const float4*>(&xs[b * K + k]);
          const float4 xb = *reinterpret_cast<const float4*>(&xs[b * K + k + 4]);
          acc[b] += wf[0] * xa.x + wf[1] * xa.y + wf[2] * xa.z + wf[3] * xa.w + wf[4] * xb.x + wf[5] * xb.y + wf[6] * xb.z +
                    wf[7] * xb.w;
        }
      }
    }
    // next row of this warp (if any) goes in flight while the reduction / epilogue of this one runs
    if (n + nwarps < N) {
      const bf16* wr = W + (long long)(n + nwarps) * K;
#pragma unroll
      for (int u = 0; u < MAXU; ++u) {
        const int k = lane * 8 + u * 256;
        wv[u] = (k < K) ? *reinterpret_cast<const uint4*>(wr + k) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      float v = acc[b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) {
        if (bias) v += bias[n];
        if (gelu) v = gelu_erf(v);
        if (residual) v += residual[(long long)b * N + n];
        if (pos) v += __bfloat162float(pos[n]);
        if (y_f32) y_f32[(long long)b * N + n] = v;
        if (y_bf16) y_bf16[(long long)b * N + n] = __float2bfloat16(v);
      }
    }
  }
}

struct SampleMeta { int off[9]; float temp[8]; float top_p[8]; };

// grid (8, B), 256 threads.  logits fp32 [B, 1280]; uniforms double [B, S, 8]; writes cur_tok[b, attr].
__global__ void __launch_bounds__(256) decode_sample_kernel(const float* __restrict__ logits, const double* __restrict__ uniforms,
                                                            const int* __restrict__ forced, const int* __restrict__ t_dev,
                                                            int* __restrict__ cur_tok, int* __restrict__ sampled, int S,
                                                            SampleMeta meta) {
  pdl_entry();
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
  // model.py:85  probs /= (sum(probs) + 1e-5)
  float part = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) part += p[i];
  const float tot1 = block_sum(part, red) + 1e-5f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = p[i] / tot1;
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
  if (threadIdx.x < 32) {
    // warp 0: lane l owns sorted entries [16 l, 16 l + 16); prefix sums by local scan + warp shuffle scan
    const int lane = threadIdx.x;
    const int k0 = lane * 16;
    float loc[16];
    float run = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { run += (k0 + i < n) ? sp[k0 + i] : 0.f; loc[i] = run; }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float up = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += up; }
    const float excl = incl - run;
    int last = 1;
    if (meta.top_p[a] < 1.0f) {            // p == 1: cumsum > 1 never holds -> top-1 (reference quirk)
      int first = 0x7fffffff;
#pragma unroll
      for (int i = 15; i >= 0; --i) if (k0 + i < n && excl + loc[i] > meta.top_p[a]) first = k0 + i;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      last = (first == 0x7fffffff) ? 1 : first + 1;
    }
    // candidate mass cs = sum of the first `last` sorted probabilities
    float csl = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) if (k0 + i < last) csl += sp[k0 + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) csl += __shfl_xor_sync(0xffffffffu, csl, o);
    const float cs = csl;
    // np.random.choice(cand, size=1, p): cdf = cumsum(p) (double), cdf /= cdf[-1], searchsorted(u, 'right')
    double dloc[16];
    double drun = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { drun += (k0 + i < last) ? (double)(sp[k0 + i] / cs) : 0.0; dloc[i] = drun; }
    double dincl = drun;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double up = __shfl_up_sync(0xffffffffu, dincl, o); if (lane >= o) dincl += up; }
    const double dexcl = dincl - drun;
    const double dtot = __shfl_sync(0xffffffffu, dincl, 31);
    const double u = uniforms[((long long)b * S + t) * 8 + a];
    int pick = 0x7fffffff;
#pragma unroll
    for (int i = 15; i >= 0; --i) if (k0 + i < last && (dexcl + dloc[i]) / dtot > u) pick = k0 + i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pick = min(pick, __shfl_xor_sync(0xffffffffu, pick, o));
    if (pick == 0x7fffffff) pick = last - 1;
    if (lane == 0) {
      const int tok = si[pick];
      sampled[((long long)b * S + t) * 8 + a] = tok;
      cur_tok[b * 8 + a] = forced ? forced[((long long)b * S + t) * 8 + a] : tok;
    }
  }
}

struct PadMeta { int pad[8]; };

// one thread per batch row: model.py:59-65
__global__ void decode_advance_kernel(const int* __restrict__ cur_tok, int* __restrict__ result, int* __restrict__ done,
                                      int* __restrict__ t_dev, int* __restrict__ n_written, int B, int S, PadMeta pm) {
  pdl_entry();
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

// Octuple2Midi truncation (reference demo.py:72-102) for a batch of generated sequences, one warp per sequence:
// the first row holding any attribute >= its <PAD> id, or a Pitch > 127 (drums are not generated), becomes the <EOS> row
// and every later row <PAD>; if there is none the LAST row becomes <EOS>.  len[b] = index of the <EOS> row (the number of
// rows handed to encoding_to_MIDI; 0 = "Generate Fail (empty)").
template <typename T>
__global__ void __launch_bounds__(128) octuple_truncate_kernel(const T* __restrict__ in, long long* __restrict__ out,
                                                              long long* __restrict__ len, int B, int S, PadMeta pm) {
  pdl_entry();
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const T* x = in + (long long)b * S * 8;
  int first = S;
  for (int i0 = 0; i0 < S && first == S; i0 += 32) {
    const int i = i0 + lane;
    bool bad = false;
    if (i < S) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long v = (long long)x[(long long)i * 8 + j];
        bad |= v >= pm.pad[j] || (j == 3 && v > 127);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if (m) first = i0 + __ffs(m) - 1;
  }
  const int eos_row = first < S ? first : S - 1;
  long long* y = out + (long long)b * S * 8;
  for (int e = lane; e < S * 8; e += 32) {
    const int i = e >> 3, j = e & 7;
    long long v = (long long)x[e];
    if (i == eos_row) v = pm.pad[j] + 3;
    else if (i > eos_row) v = pm.pad[j];
    y[e] = v;
  }
  if (lane == 0) len[b] = eos_row;
}

}  // namespace

#define PB_STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int pb_decode_finalize(float* acc, const float* bias, const void* residual, const void* pos_table, const int* t_dev,
                                  const float* gamma, const float* beta, void* out, float* out_f32, int B, int N, int gelu,
                                  void* stream) {
  if (N * 4 > 48 * 1024) return pb_set_error("decode_finalize: row too wide");
  PB_LAUNCH((decode_finalize_kernel), B, 256, N * sizeof(float), PB_STREAM(stream), acc, bias, (const bf16*)residual, (const bf16*)pos_table,
                                                                          t_dev, gamma, beta, (bf16*)out, out_f32, N, gelu, 1e-5f);
  return pb_check_launch("decode_finalize");
}

template <int B>
static void gemv_launch(const float* x_raw, const float* gamma, const float* beta, float* x_norm_out, const void* W,
                        const float* bias, const float* residual, const void* pos_table, const int* t_dev, float* y_f32,
                        void* y_bf16, int N, int K, int gelu, cudaStream_t st) {
  const int smem = B * K * (int)sizeof(float);
  static bool attr = false;
  if (!attr && smem > 48 * 1024) {
    cudaFuncSetAttribute(decode_gemv_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2048 * 4);
    attr = true;
  }
  int grid = (N + 7) / 8;                       // 8 warps per CTA, one weight row per warp per pass
  const int cap = pb_num_sms() * 4;
  if (grid > cap) grid = cap;
  PB_LAUNCH((decode_gemv_kernel<B>), grid, 256, smem, st, x_raw, gamma, beta, x_norm_out, (const bf16*)W, bias, residual,
                                                 (const bf16*)pos_table, t_dev, y_f32, (bf16*)y_bf16, N, K, gelu, 1e-5f);
}

extern "C" int pb_decode_gemv(const float* x_raw, const float* gamma, const float* beta, float* x_norm_out, const void* W,
                              const float* bias, const float* residual, const void* pos_table, const int* t_dev, float* y_f32,
                              void* y_bf16, int B, int N, int K, int gelu, void* stream) {
  if (B < 1 || B > GV_MAXB) return pb_set_error("decode_gemv: batch must be in [1, 8]");
  if (K % 8 != 0 || K > 2048) return pb_set_error("decode_gemv: K must be a multiple of 8 and <= 2048");
  cudaStream_t st = PB_STREAM(stream);
  switch (B) {
    case 1: gemv_launch<1>(x_raw, gamma, beta, x_norm_out, W, bias, residual, pos_table, t_dev, y_f32, y_bf16, N, K, gelu, st); break;
    case 2: gemv_launch<2>(x_raw, gamma, beta, x_norm_out, W, bias, residual, pos_table, t_dev, y_f32, y_bf16, N, K, gelu, st); break;
    case 3: gemv_launch<3>(x_raw, gamma, beta, x_norm_out, W, bias, residual, pos_table, t_dev, y_f32, y_bf16, N, K, gelu, st); break;
    case 4: gemv_launch<4>(x_raw, gamma, beta, x_norm_out, W, bias, residual, pos_table, t_dev, y_f32, y_bf16, N, K, gelu, st); break;
    case 5: gemv_launch<5>(x_raw, gamma, beta, x_norm_out, W, bias, residual, pos_table, t_dev, y_f32, y_bf16, N, K, gelu, st); break;
    case 6: gemv_launch<6>(x_raw, gamma, beta, x_norm_out, W, bias, residual, pos_table, t_dev, y_f32, y_bf16, N, K, gelu, st); break;
    case 7: gemv_launch<7>(x_raw, gamma, beta, x_norm_out, W, bias, residual, pos_table, t_dev, y_f32, y_bf16, N, K, gelu, st); break;
    default: gemv_launch<8>(x_raw, gamma, beta, x_norm_out, W, bias, residual, pos_table, t_dev, y_f32, y_bf16, N, K, gelu, st); break;
  }
  return pb_check_launch("decode_gemv");
}

extern "C" int pb_decode_attn(const void* q, int q_ld, const void* k_new, const void* v_new, void* k_cache, void* v_cache,
                              long long kv_batch_stride, int kv_ld, const uint8_t* key_keep, int n_keys, const int* t_dev,
                              int append, void* out, int out_ld, int B, int H, int hd, float scale, int max_keys,
                              float* workspace, int* tickets, float* out_f32, void* stream) {
  if (hd > 128 || (hd % 8) != 0) return pb_set_error("decode_attn: head_dim must be a multiple of 8 and <= 128");
  if ((q_ld % 4) != 0 || (kv_ld % 4) != 0) return pb_set_error("decode_attn: row strides must be multiples of 4 elements");
  const int NS = (max_keys + DK - 1) / DK;
  dim3 grid(H, B, NS);
  PB_LAUNCH((decode_attn_kernel), grid, 128, 0, PB_STREAM(stream), 
      (const bf16*)q, q_ld, (const bf16*)k_new, (const bf16*)v_new, (bf16*)k_cache, (bf16*)v_cache, kv_batch_stride, kv_ld, key_keep,
      n_keys, t_dev, append, (bf16*)out, out_ld, hd, scale, max_keys, workspace, tickets, out_f32);
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
  PB_LAUNCH((decode_sample_kernel), grid, 256, 0, PB_STREAM(stream), logits, uniforms, forced, t_dev, cur_tok, sampled, S, m);
  return pb_check_launch("decode_sample");
}

extern "C" int pb_decode_advance(const int* cur_tok, int* result, int* done, int* t_dev, int* n_written, int B, int S,
                                 const int* pad_host, void* stream) {
  if (B > 1024) return pb_set_error("decode_advance: batch > 1024");
  PadMeta pm;
  for (int i = 0; i < 8; ++i) pm.pad[i] = pad_host[i];
  PB_LAUNCH((decode_advance_kernel), 1, ((B + 31) / 32) * 32, 0, PB_STREAM(stream), cur_tok, result, done, t_dev, n_written, B, S, pm);
  return pb_check_launch("decode_advance");
}

extern "C" int pb_octuple_truncate(const void* ids, int ids_int64, long long* out, long long* len, int B, int S,
                                   const int* pad_host, void* stream) {
  if (B <= 0 || S <= 0) return pb_set_error("octuple_truncate: empty batch");
  PadMeta pm;
  for (int i = 0; i < 8; ++i) pm.pad[i] = pad_host[i];
  const int grid = (B + 3) / 4;
  if (ids_int64) {
    PB_LAUNCH((octuple_truncate_kernel<long long>), grid, 128, 0, PB_STREAM(stream), (const long long*)ids, out, len, B, S, pm);
  } else {
    PB_LAUNCH((octuple_truncate_kernel<int>), grid, 128, 0, PB_STREAM(stream), (const int*)ids, out, len, B, S, pm);
  }
  return pb_check_launch("octuple_truncate");
}

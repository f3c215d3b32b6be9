// Persistent whole-step decode kernel for batch 1 (reference model.py:28-107: PianoBartLM.forward(generate=True) +
// sample / sampling / nucleus), default model geometry (d = 1024, 8 heads x 128, ffn 2048, Octuple vocab 1280).
//
// The reference re-runs encoder + 1024-position decoder per generated token.  Round 1 replaced that by a KV cache and one
// CUDA-graph replay of 70 small kernels per token, which was bound by ~8 us of launch latency per dependent kernel
// (5.9 % of the HBM roofline).  This kernel runs N tokens in ONE launch:
//
//   * one CTA per SM (cooperative launch: all CTAs are co-resident), 16 consumer warps + 1 producer warp;
//   * WEIGHTS AND KV-CACHE SLICES NEVER WAIT FOR ACTIVATIONS: the producer thread streams this CTA's slice of every
//     weight matrix and of the K/V caches through an 8-slot shared-memory ring with 1-D bulk copies (cp.async.bulk +
//     mbarrier complete_tx), running up to 8 chunks (~1 decoder layer) ahead of the consumers, across layer and token
//     boundaries - HBM streaming is decoupled from the dependency chain;
//   * the dependency chain (83 "hops" per token: 6 projections + 2 x (attention partial, combine) per layer, front end,
//     heads, sampler) exchanges activation vectors through global memory WITHOUT grid barriers: every 8-byte word carries
//     its payload (two bf16 or one fp32) and a 32-bit tag that is unique per (token, hop); consumers poll the words they
//     need until the tag matches (single-copy-atomic 64-bit accesses, the "LL" protocol of collective libraries).  A hop
//     costs one L2 round trip instead of a release fence + atomic + acquire spin;
//   * LayerNorm, bias, GELU, residual and the position row are computed in registers / shared memory by the consumer of
//     the vector; the sampler (temperature softmax, nucleus, numpy's choice semantics) is the last hop.
//
// Work split.  Projection y = W x with N output rows: CTA c owns the row PAIRS [c P / G, (c+1) P / G), P = N / 2 (a pair
// = one exchanged word).  Attention: CTA c < 144 owns (head c / 18, key split c % 18); key j of a sequence lives in split
// j % 18, slot j / 18 of the split-major cache layout [layer][head][split][slot][128], so a CTA's keys are contiguous in
// memory (one bulk copy) and balanced for every cache length.  Partials (max, sum, out[128]) of the 18 splits of a head
// are combined by the first CTA of the head.
#include "decode_common.cuh"

namespace pbdec {

// Activation vectors live in shared memory as fp32 in a permuted order so that a lane's 8 consecutive elements (the
// elements its 16-byte weight load covers) are two conflict-free float4 loads: element e -> chunk e / 256, lane (e % 256) / 8.
__device__ __forceinline__ int xperm(int e) {
  const int r = e & 255;
  return (e & ~255) + ((r & 4) << 5) + ((r >> 3) << 2) + (r & 3);
}

// Exchange vectors that EVERY CTA polls (raw0/1/2, qkv, qc, ob, f1) keep only 4 words (one 32-byte sector) per 128-byte
// line: the polling traffic of 148 SMs then spreads over 4x more L2 slices (measured: with dense vectors the ~32 hot lines
// of a vector made an L2 round trip cost 700-2500 cycles instead of ~300, and that latency is paid twice per hop).
__device__ __forceinline__ int lls(int w) { return ((w >> 2) << 4) | (w & 3); }

// pair range of CTA c for a projection with P row pairs
__device__ __forceinline__ void pair_range(int P, int c, int G, int& p0, int& p1) {
  p0 = (int)(((unsigned)c * (unsigned)P) / (unsigned)G);            // c < 2^8, P <= 1536: 32-bit arithmetic is exact
  p1 = (int)(((unsigned)(c + 1) * (unsigned)P) / (unsigned)G);
}
// cached keys of split s when the sequence holds `n` keys (keys j < n with j % NSPLIT == s)
__device__ __forceinline__ int split_count(int n, int s) { return n > s ? (n - s + NSPLIT - 1) / NSPLIT : 0; }

// ------------------------------------------------------------------------------------------------ chunk schedule
// Per token, in consumption order: in_linear (2 chunks); per layer: qkv (2), self K, self V, out_proj, q_c, cross K,
// cross V, out_proj_c, fc1 (2), fc2 (2); heads (1).  Producer and consumers enumerate the same list.
struct Chunk { const void* src; uint32_t bytes; int kv; };

struct Sched {
  const pb_decode_persist_desc* p;
  int c, G, h, s;       // CTA index, grid size, attention head / split of this CTA (h < 0: no attention work)
  __device__ __forceinline__ Chunk proj(const void* w, int N, int K, int half, int nhalves) const {
    int p0, p1;
    pair_range(N / 2, c, G, p0, p1);
    int a = p0, b = p1;
    if (nhalves == 2) {
      const int mid = p0 + (p1 - p0 + 1) / 2;
      if (half == 0) b = mid; else a = mid;
    }
    Chunk ch;
    ch.src = reinterpret_cast<const bf16*>(w) + (long long)(2 * a) * K;
    ch.bytes = (uint32_t)((b - a) * 2 * K * 2);
    ch.kv = 0;
    return ch;
  }
  __device__ __forceinline__ Chunk kv(const void* base, int nslots) const {
    Chunk ch;
    ch.kv = 1;
    if (h < 0 || nslots <= 0) { ch.src = base; ch.bytes = 0; return ch; }
    ch.src = reinterpret_cast<const bf16*>(base) + ((long long)(h * NSPLIT + s) * NSLOT) * HD;
    ch.bytes = (uint32_t)(nslots * HD * 2);
    return ch;
  }
  // chunk `i` (0 .. chunks_per_token-1) of the token whose cache length before the step is t
  __device__ __forceinline__ Chunk get(int i, int t) const {
    if (i < 2) return proj(p->w_in, D, E, i, 2);
    i -= 2;
    const int l = i / 14, k = i % 14;
    if (l >= p->n_layers) return proj(p->w_heads, V, D, 0, 1);
    const pb_decode_layer& L = p->layer[l];
    switch (k) {
      case 0: return proj(L.wqkv, 3 * D, D, 0, 2);
      case 1: return proj(L.wqkv, 3 * D, D, 1, 2);
      case 2: return kv(L.self_k, h < 0 ? 0 : split_count(t, s));
      case 3: return kv(L.self_v, h < 0 ? 0 : split_count(t, s));
      case 4: return proj(L.wo, D, D, 0, 1);
      case 5: return proj(L.wqc, D, D, 0, 1);
      case 6: return kv(L.cross_k, h < 0 ? 0 : split_count(p->S_enc, s));
      case 7: return kv(L.cross_v, h < 0 ? 0 : split_count(p->S_enc, s));
      case 8: return proj(L.woc, D, D, 0, 1);
      case 9: return proj(L.w1, F, D, 0, 2);
      case 10: return proj(L.w1, F, D, 1, 2);
      case 11: return proj(L.w2, D, F, 0, 2);
      case 12: return proj(L.w2, D, F, 1, 2);
      default: return proj(L.w2, D, F, 1, 2);   // (k == 13 is not used: see chunks_per_token)
    }
  }
};
// 13 chunks per layer are enumerated with stride 14 to keep the index arithmetic a shift-free div; slot 13 is skipped
__device__ __forceinline__ bool chunk_used(int i, int n_layers) {
  if (i < 2) return true;
  const int k = (i - 2) % 14, l = (i - 2) / 14;
  if (l >= n_layers) return (i - 2) == 14 * n_layers;
  return k != 13;
}

struct Shared {
  alignas(1024) uint8_t ring[RING][SLOT_BYTES];
  alignas(16) float x[2 * 1024];        // projection input (permuted order), up to 2048 elements
  alignas(16) float res[1024];          // normalised residual stream (permuted order)
  float out[32];                        // per-row results of the current projection (rows of this CTA)
  float q[HD], knew[HD], vnew[HD];
  float sc[64];                         // scores / probabilities of this CTA's keys (+ the new key)
  float po[8][HD];
  float red[2 * NCW];
  float stat[4];
  uint8_t ckeep[64];                    // encoder key-padding flags of this CTA's cross-attention slots
  float sp[512]; float sprob[512]; int sidx[512];     // sampler scratch
  int tok[8];
  int stop;
  alignas(8) uint64_t full_bar[RING];
  alignas(8) uint64_t empty_bar[RING];
};

// ------------------------------------------------------------------------------------------------ consumer state
struct Ctx {
  const pb_decode_persist_desc* p;
  Shared* sm;
  int c, G, h, s, tid, warp, lane;
  uint32_t chunk_no;     // running index of the next chunk to consume (ring slot = chunk_no % RING)
  int* err;
  long long* tr;         // developer trace (thread 0 of the CTA, last token of the launch) or null
  int tr_n;
  __device__ __forceinline__ void stamp() {
    if (tr != nullptr && tid == 0 && tr_n < 6 * 96) tr[tr_n++] = clock64();
  }
  __device__ __forceinline__ const uint8_t* chunk_wait() {
    const uint32_t slot = chunk_no % RING;
    mbar_wait_to(&sm->full_bar[slot], (chunk_no / RING) & 1u, err);
    return sm->ring[slot];
  }
  // all consumer threads are done with the chunk (caller has synchronised them): hand the slot back
  __device__ __forceinline__ void chunk_release() {
    if (tid == 0) pb::mbar_arrive(&sm->empty_bar[chunk_no % RING]);
    ++chunk_no;
  }
};

// reads a bf16x2-tagged vector of n elements (n / 2 words) into shared memory (permuted fp32)
__device__ __forceinline__ void read_vec(Ctx& cx, const unsigned long long* buf, int n, uint32_t tag, float* dst) {
  cx.stamp();                            // [6h+0] hop start
  for (int w = cx.tid; w < n / 2; w += NCONS) {
    const float2 f = unpack_bf16x2(ll_wait(buf + lls(w), tag, cx.err));
    dst[xperm(2 * w)] = f.x;
    dst[xperm(2 * w + 1)] = f.y;
  }
  cx.stamp();                            // [6h+1] this thread's words arrived
  cons_sync();
}
// sum and sum of squares over the consumer threads in one pass (two named-barrier syncs)
__device__ __forceinline__ float2 cons_sum2(float a, float b, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  cons_sync();
  if (l == 0) { red[w] = a; red[NCW + w] = b; }
  cons_sync();
  float ta = 0.f, tb = 0.f;
#pragma unroll
  for (int i = 0; i < NCW; ++i) { ta += red[i]; tb += red[NCW + i]; }
  return make_float2(ta, tb);
}
// dst (permuted fp32, D elements) = LayerNorm(tagged vector `buf`): every thread owns one word (two elements); gamma / beta
// are requested before the poll so that their latency hides behind the wait for the data
__device__ __forceinline__ void read_ln(Ctx& cx, const unsigned long long* buf, uint32_t tag, const float* gamma,
                                        const float* beta, float* dst) {
  cx.stamp();                            // [6h+0] hop start
  const float2 g = *reinterpret_cast<const float2*>(gamma + 2 * cx.tid);
  const float2 b = *reinterpret_cast<const float2*>(beta + 2 * cx.tid);
  const float2 f = unpack_bf16x2(ll_wait(buf + lls(cx.tid), tag, cx.err));
  cx.stamp();                            // [6h+1] this thread's word arrived
  const float2 st = cons_sum2(f.x + f.y, f.x * f.x + f.y * f.y, cx.sm->red);
  const float mean = st.x * (1.0f / D);
  const float var = fmaxf(st.y * (1.0f / D) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  dst[xperm(2 * cx.tid)] = (f.x - mean) * rstd * g.x + b.x;
  dst[xperm(2 * cx.tid + 1)] = (f.y - mean) * rstd * g.y + b.y;
  cons_sync();
}

// One projection hop: rows [2 p0, 2 p1) of W (K columns, staged in 1 or 2 ring chunks) times the vector `xin` (permuted
// fp32 in shared memory); epilogue bias / GELU / residual (permuted smem vector) / position row; result written as tagged
// bf16x2 words (out_bf) or tagged fp32 words (out_f32).
template <int K>
__device__ __forceinline__ void proj_hop(Ctx& cx, int N, int nchunks, const float* xin, const float* bias, const float* resid,
                                         const bf16* pos_row, bool gelu, unsigned long long* out_bf,
                                         unsigned long long* out_f32, uint32_t tag) {
  int p0, p1;
  pair_range(N / 2, cx.c, cx.G, p0, p1);
  const int nrows = 2 * (p1 - p0);
  const int mid_rows = nchunks == 2 ? 2 * ((p1 - p0 + 1) / 2) : nrows;
  // epilogue operands of the rows this thread will finish: requested now, consumed after the dot products
  float eb0 = 0.f, eb1 = 0.f;
  if (cx.tid < nrows / 2) {
    const int n0 = 2 * (p0 + cx.tid);
    if (bias) { eb0 = bias[n0]; eb1 = bias[n0 + 1]; }
    if (pos_row) { eb0 += __bfloat162float(pos_row[n0]); eb1 += __bfloat162float(pos_row[n0 + 1]); }
  }
  const uint8_t* base0 = nullptr;
  const uint8_t* base1 = nullptr;
  cx.stamp();                            // [6h+2] input vector complete (hop h)
  base0 = cx.chunk_wait();
  uint32_t c0 = cx.chunk_no;
  if (nchunks == 2) { ++cx.chunk_no; base1 = cx.chunk_wait(); cx.chunk_no = c0; }
  cx.stamp();                            // [6h+3] weights in shared memory
  // each warp takes rows warp, warp + 16 (at most two: nrows <= 22)
  float acc[2] = {0.f, 0.f};
  const uint8_t* wrow[2];
  bool have[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int r = cx.warp + u * NCW;
    have[u] = r < nrows;
    const int rr = have[u] ? r : 0;
    wrow[u] = (rr < mid_rows ? base0 + (size_t)rr * K * 2 : base1 + (size_t)(rr - mid_rows) * K * 2) + cx.lane * 16;
  }
  if (have[0]) {
#pragma unroll
    for (int j = 0; j < K / 256; ++j) {
      const float4 xa = *reinterpret_cast<const float4*>(xin + j * 256 + cx.lane * 4);
      const float4 xb = *reinterpret_cast<const float4*>(xin + j * 256 + 128 + cx.lane * 4);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !have[1]) break;
        const uint4 wv = *reinterpret_cast<const uint4*>(wrow[u] + j * 512);
        const float2 w0 = unpack_bf16x2(wv.x), w1 = unpack_bf16x2(wv.y), w2 = unpack_bf16x2(wv.z), w3 = unpack_bf16x2(wv.w);
        acc[u] += w0.x * xa.x + w0.y * xa.y + w1.x * xa.z + w1.y * xa.w + w2.x * xb.x + w2.y * xb.y + w3.x * xb.z + w3.y * xb.w;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float v = acc[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (cx.lane == 0 && have[u]) cx.sm->out[cx.warp + u * NCW] = v;
    }
  }
  cons_sync();                           // results published; every warp is done with the weight chunks
  cx.stamp();                            // [6h+4] dot products done
  cx.chunk_release();
  if (nchunks == 2) cx.chunk_release();
  if (cx.tid < nrows / 2) {
    const int n0 = 2 * (p0 + cx.tid);
    float v0 = cx.sm->out[2 * cx.tid] + eb0, v1 = cx.sm->out[2 * cx.tid + 1] + eb1;   // (the position row is only used without GELU)
    if (gelu) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
    if (resid) { v0 += resid[xperm(n0)]; v1 += resid[xperm(n0 + 1)]; }
    if (out_bf) ll_store(out_bf + lls(n0 >> 1), pack_bf16x2(v0, v1), tag);
    if (out_f32) { ll_store(out_f32 + n0, __float_as_uint(v0), tag); ll_store(out_f32 + n0 + 1, __float_as_uint(v1), tag); }
  }
  cx.stamp();                            // [6h+5] results stored
}

// Attention partial of (head h, split s): `nold` cached keys in the K / V ring chunks (+ the new key of this step when
// `own_new`), optional key-padding flags; writes (max, sum, out[128]) as tagged fp32 words.
__device__ __forceinline__ void attn_partial(Ctx& cx, int nold, bool own_new, bool use_keep, unsigned long long* part,
                                             uint32_t tag) {
  Shared* sm = cx.sm;
  cx.stamp();
  const uint8_t* kc = cx.chunk_wait();
  const uint32_t c0 = cx.chunk_no;
  ++cx.chunk_no;
  const uint8_t* vc = cx.chunk_wait();
  cx.chunk_no = c0;
  cx.stamp();
  // scores: warp w takes keys w, w + 16, ...; lane = 4 head dims
  const float4 q4 = *reinterpret_cast<const float4*>(&sm->q[cx.lane * 4]);
  for (int j = cx.warp; j < nold; j += NCW) {
    const uint2 kw = *reinterpret_cast<const uint2*>(kc + (size_t)j * (HD * 2) + cx.lane * 8);
    const float2 k0 = unpack_bf16x2(kw.x), k1 = unpack_bf16x2(kw.y);
    float d = q4.x * k0.x + q4.y * k0.y + q4.z * k1.x + q4.w * k1.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (cx.lane == 0) sm->sc[j] = (use_keep && !sm->ckeep[j]) ? -INFINITY : d;
  }
  if (own_new && cx.warp == NCW - 1) {
    const float4 k4 = *reinterpret_cast<const float4*>(&sm->knew[cx.lane * 4]);
    float d = q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (cx.lane == 0) sm->sc[nold] = d;
  }
  cons_sync();
  const int nk = nold + (own_new ? 1 : 0);
  if (cx.warp == 0) {
    float m = -INFINITY;
    for (int j = cx.lane; j < nk; j += 32) m = fmaxf(m, sm->sc[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int j = cx.lane; j < nk; j += 32) {
      const float e = (m == -INFINITY) ? 0.f : __expf(sm->sc[j] - m);
      sm->sc[j] = e;
      l += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (cx.lane == 0) { sm->stat[0] = m; sm->stat[1] = l; }
  }
  cons_sync();
  // P V: thread = (key group g of 8, dim pair dp of 64)
  {
    const int g = cx.tid >> 6, dp = cx.tid & 63;
    float a0 = 0.f, a1 = 0.f;
    for (int j = g; j < nold; j += 8) {
      const float2 v2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vc + (size_t)j * (HD * 2) + dp * 4));
      const float pj = sm->sc[j];
      a0 += pj * v2.x;
      a1 += pj * v2.y;
    }
    sm->po[g][2 * dp] = a0;
    sm->po[g][2 * dp + 1] = a1;
  }
  cons_sync();
  cx.stamp();
  cx.chunk_release();
  cx.chunk_release();
  if (cx.tid < HD) {
    float o = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) o += sm->po[g][cx.tid];
    if (own_new) o += sm->sc[nold] * sm->vnew[cx.tid];
    ll_store(part + 4 + cx.tid, __float_as_uint(o), tag);
  } else if (cx.tid < HD + 2) {
    ll_store(part + (cx.tid - HD), __float_as_uint(sm->stat[cx.tid - HD]), tag);
  }
  cx.stamp();
}

// combine the NSPLIT partials of head h -> o[h*128 .. +128) as tagged bf16x2 words
__device__ __forceinline__ void attn_combine(Ctx& cx, int h, const unsigned long long* part_h, uint32_t tag_in,
                                             unsigned long long* obuf, uint32_t tag_out) {
  Shared* sm = cx.sm;
  if (cx.tid < 2 * NSPLIT) {
    const int s = cx.tid >> 1, which = cx.tid & 1;
    sm->po[0][cx.tid] = __uint_as_float(ll_wait(part_h + (size_t)s * PB_DECODE_PART_WORDS + which, tag_in, cx.err));
  }
  cons_sync();
  if (cx.tid < HD) {
    float M = -INFINITY;
#pragma unroll
    for (int s = 0; s < NSPLIT; ++s) M = fmaxf(M, sm->po[0][2 * s]);
    float Lsum = 0.f, acc = 0.f;
    // all 18 partial words of this dim are requested before the first one is examined (18 dependent L2 round trips made
    // this hop the longest of the layer in the first version); a word whose tag is not there yet is re-polled
    unsigned long long pw[NSPLIT];
#pragma unroll
    for (int s = 0; s < NSPLIT; ++s) pw[s] = ll_load(part_h + (size_t)s * PB_DECODE_PART_WORDS + 4 + cx.tid);
#pragma unroll
    for (int s = 0; s < NSPLIT; ++s) {
      const float m = sm->po[0][2 * s];
      const float w = (m == -INFINITY) ? 0.f : __expf(m - M);
      uint32_t ow = (uint32_t)pw[s];
      if ((uint32_t)(pw[s] >> 32) != tag_in) ow = ll_wait(part_h + (size_t)s * PB_DECODE_PART_WORDS + 4 + cx.tid, tag_in, cx.err);
      Lsum += w * sm->po[0][2 * s + 1];
      acc += w * __uint_as_float(ow);
    }
    const float r = Lsum > 0.f ? acc / Lsum : 0.f;
    const float other = __shfl_xor_sync(0xffffffffu, r, 1);
    if ((cx.tid & 1) == 0) ll_store(obuf + lls((h * HD + cx.tid) >> 1), pack_bf16x2(r, other), tag_out);
  }
  cons_sync();
}

// sampler of one attribute: see sample_core (decode_common.cuh); logits in sm->sp
__device__ __forceinline__ int sample_attr(Ctx& cx, int n, float temp, float top_p, double u) {
  return sample_core(cx.tid, n, temp, top_p, u, cx.sm->sp, cx.sm->sprob, cx.sm->sidx, cx.sm->red, &cx.sm->tok[0]);
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NTHREADS, 1) decode_persist_kernel(const __grid_constant__ pb_decode_persist_desc P,
                                                                       const SampleMeta meta, int n_steps) {
  extern __shared__ uint8_t smem_raw[];
  Shared* sm = reinterpret_cast<Shared*>(smem_raw + ((1024u - (pb::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x, G = gridDim.x;
  const int nl = P.n_layers;
  int* err = P.error_flag;
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) { pb::mbar_init(&sm->full_bar[i], 1); pb::mbar_init(&sm->empty_bar[i], 1); }
    pb::fence_mbar_init();
    sm->stop = 0;
  }
  __syncthreads();
  const int t0 = *P.t_dev;                       // cache length at entry (same value in every CTA: written by the previous launch)
  n_steps = max(0, min(n_steps, P.S_max - t0));  // never decode past the cache / result capacity
  const uint32_t epoch0 = *P.epoch;              // tag base; advanced by CTA 0 at the end of the launch
  const int att_h = c < H * NSPLIT ? c / NSPLIT : -1, att_s = c % NSPLIT;
  const int chunks_per_token = 2 + 14 * nl + 1;

  if (warp == NCW) {
    // ============================================================ producer: weight / KV streaming
    if (lane == 0) {
      Sched sch{&P, c, G, att_h, att_s};
      const uint64_t pol_w = l2_policy_evict_first(), pol_kv = l2_policy_evict_last();
      uint32_t k = 0;
      bool stopped = false;
      for (int st = 0; st < n_steps && !stopped; ++st) {
        const int t = t0 + st;
        for (int i = 0; i < chunks_per_token; ++i) {
          if (!chunk_used(i, nl)) continue;
          const uint32_t slot = k % RING;
          // wait until the consumers released the chunk that used this slot RING chunks ago (or a stop request)
          if (k >= RING) {
            const uint32_t par = ((k / RING) & 1u) ^ 1u;
            const long long w0 = clock64();
            uint32_t n = 0;
            while (!pb::mbar_try_wait(&sm->empty_bar[slot], par)) {
              if (*reinterpret_cast<volatile int*>(&sm->stop)) { stopped = true; break; }
              if ((++n & 255u) == 0 && clock64() - w0 > TIMEOUT_CYCLES) die(err, 4);
            }
            if (stopped) break;
          }
          const Chunk ch = sch.get(i, t);
          if (ch.bytes > 0 && !(P.dbg_flags & 1)) {
            pb::mbar_expect_tx(&sm->full_bar[slot], ch.bytes);
            bulk_load(sm->ring[slot], ch.src, ch.bytes, &sm->full_bar[slot], ch.kv ? pol_kv : pol_w);
          } else {
            pb::mbar_arrive(&sm->full_bar[slot]);
          }
          ++k;
        }
      }
      // drain: every bulk copy issued by this thread has landed before the CTA may exit.  Chunks the consumers did not
      // take (early stop) are the last min(RING, ...) ones; waiting for the current phase of every slot's full barrier
      // covers them (a phase that was already consumed returns immediately).
      if (stopped) {
        for (uint32_t j = (k > RING ? k - RING : 0); j < k; ++j) mbar_wait_to(&sm->full_bar[j % RING], (j / RING) & 1u, err);
      }
    }
    return;
  }

  // ============================================================== consumers
  Ctx cx;
  cx.p = &P; cx.sm = sm; cx.c = c; cx.G = G; cx.h = att_h; cx.s = att_s; cx.tid = tid; cx.warp = warp; cx.lane = lane;
  cx.chunk_no = 0; cx.err = err; cx.tr = nullptr; cx.tr_n = 0;
  // encoder key-padding flags of this CTA's cross-attention slots (static for the whole generation)
  if (tid < 64) {
    const int j = tid * NSPLIT + att_s;
    sm->ckeep[tid] = (att_h >= 0 && j < P.S_enc && (P.enc_keep == nullptr || P.enc_keep[j] != 0)) ? 1 : 0;
  }
  int cur[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) cur[a] = P.cur_tok[a];
  bool done = P.done[0] != 0;
  cons_sync();

  int steps_run = 0;
  for (int st = 0; st < n_steps; ++st) {
    const int t = t0 + st;
    const uint32_t tb = epoch0 + (uint32_t)st * HOPS + 1u;       // tag of hop i of this token = tb + i
    uint32_t hop = 0;
    cx.tr = (P.trace != nullptr && st == n_steps - 1) ? P.trace + (size_t)c * (6 * 96) : nullptr;
    cx.stamp();
    // ---------------- front end: 8 embedding rows (table pre-scaled by 16, PianoBart.py:9-16,60-67) -> in_linear + pos
    for (int e2 = tid; e2 < E / 2; e2 += NCONS) {
      const int a = e2 >> 7, col = (e2 & 127) * 2;
      const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const bf16*>(P.emb_table) +
                                                            (size_t)(meta.off[a] + cur[a]) * 256 + col);
      const float2 f = unpack_bf16x2(w);
      sm->x[xperm(2 * e2)] = f.x;
      sm->x[xperm(2 * e2 + 1)] = f.y;
    }
    cx.stamp();
    cons_sync();
    proj_hop<E>(cx, D, 2, sm->x, P.b_in, nullptr, reinterpret_cast<const bf16*>(P.pos_table) + (size_t)(t + 2) * D, false,
                P.raw0, nullptr, tb + hop);
    const float* ln_g = P.lne_g; const float* ln_b = P.lne_b;
    for (int l = 0; l < nl; ++l) {
      const pb_decode_layer& L = P.layer[l];
      // ---- h = LN(raw0) ; qkv
      read_ln(cx, P.raw0, tb + hop, ln_g, ln_b, sm->res);
      ++hop;
      proj_hop<D>(cx, 3 * D, 2, sm->res, L.bqkv, nullptr, nullptr, false, P.qkv, nullptr, tb + hop);
      // ---- self attention partial
      const uint32_t tag_qkv = tb + hop;
      ++hop;
      if (att_h >= 0) {
        const bool own_new = (t % NSPLIT) == att_s;
        cx.stamp();
        if (tid < 3 * (HD / 2)) {
          const int which = tid / (HD / 2), w = tid % (HD / 2);
          const float2 f = unpack_bf16x2(ll_wait(P.qkv + lls((which * D + att_h * HD) / 2 + w), tag_qkv, err));
          float* dst = which == 0 ? sm->q : (which == 1 ? sm->knew : sm->vnew);
          const float sc = which == 0 ? 0.08838834764831845f : 1.0f;       // hd^-0.5
          dst[2 * w] = f.x * sc;
          dst[2 * w + 1] = f.y * sc;
          if (own_new && which > 0) {
            // append: key t -> split t % 18 (this CTA), slot t / 18 of the split-major cache
            bf16* dstc = reinterpret_cast<bf16*>(which == 1 ? L.self_k : L.self_v) +
                         ((size_t)(att_h * NSPLIT + att_s) * NSLOT + t / NSPLIT) * HD;
            *reinterpret_cast<uint32_t*>(dstc + 2 * w) = pack_bf16x2(f.x, f.y);
          }
        }
        if (own_new) asm volatile("fence.proxy.async.global;" ::: "memory");   // later bulk copies (async proxy) read these rows
        cx.stamp();
        cons_sync();
        attn_partial(cx, split_count(t, att_s), own_new, false,
                     P.part + (size_t)(att_h * NSPLIT + att_s) * PB_DECODE_PART_WORDS, tb + hop);
      } else {
        // no attention work: still walk the (empty) K / V chunks so that the ring stays in step with the producer
        cx.chunk_wait(); cons_sync(); cx.chunk_release();
        cx.chunk_wait(); cons_sync(); cx.chunk_release();
      }
      const uint32_t tag_part = tb + hop;
      ++hop;
      if (att_h >= 0 && att_s == 0) attn_combine(cx, att_h, P.part + (size_t)(att_h * NSPLIT) * PB_DECODE_PART_WORDS, tag_part, P.ob, tb + hop);
      // ---- out_proj + residual h
      read_vec(cx, P.ob, D, tb + hop, sm->x);
      ++hop;
      proj_hop<D>(cx, D, 1, sm->x, L.bo, sm->res, nullptr, false, P.raw1, nullptr, tb + hop);
      // ---- h1 = LN1 ; q_c
      read_ln(cx, P.raw1, tb + hop, L.ln1_g, L.ln1_b, sm->res);
      ++hop;
      proj_hop<D>(cx, D, 1, sm->res, L.bqc, nullptr, nullptr, false, P.qc, nullptr, tb + hop);
      const uint32_t tag_qc = tb + hop;
      ++hop;
      // ---- cross attention partial
      if (att_h >= 0) {
        cx.stamp();
        if (tid < HD / 2) {
          const float2 f = unpack_bf16x2(ll_wait(P.qc + lls((att_h * HD) / 2 + tid), tag_qc, err));
          sm->q[2 * tid] = f.x * 0.08838834764831845f;
          sm->q[2 * tid + 1] = f.y * 0.08838834764831845f;
        }
        cx.stamp();
        cons_sync();
        attn_partial(cx, split_count(P.S_enc, att_s), false, true,
                     P.part + (size_t)(att_h * NSPLIT + att_s) * PB_DECODE_PART_WORDS, tb + hop);
      } else {
        cx.chunk_wait(); cons_sync(); cx.chunk_release();
        cx.chunk_wait(); cons_sync(); cx.chunk_release();
      }
      const uint32_t tag_cpart = tb + hop;
      ++hop;
      if (att_h >= 0 && att_s == 0) attn_combine(cx, att_h, P.part + (size_t)(att_h * NSPLIT) * PB_DECODE_PART_WORDS, tag_cpart, P.ob, tb + hop);
      read_vec(cx, P.ob, D, tb + hop, sm->x);
      ++hop;
      proj_hop<D>(cx, D, 1, sm->x, L.boc, sm->res, nullptr, false, P.raw2, nullptr, tb + hop);
      // ---- h2 = LN2 ; fc1 + GELU ; fc2 + residual h2
      read_ln(cx, P.raw2, tb + hop, L.ln2_g, L.ln2_b, sm->res);
      ++hop;
      proj_hop<D>(cx, F, 2, sm->res, L.b1, nullptr, nullptr, true, P.f1, nullptr, tb + hop);
      read_vec(cx, P.f1, F, tb + hop, sm->x);
      ++hop;
      proj_hop<F>(cx, D, 2, sm->x, L.b2, sm->res, nullptr, false, P.raw0, nullptr, tb + hop);
      ln_g = L.ln3_g; ln_b = L.ln3_b;
    }
    // ---------------- heads
    read_ln(cx, P.raw0, tb + hop, ln_g, ln_b, sm->res);
    ++hop;
    proj_hop<D>(cx, V, 1, sm->res, P.b_heads, nullptr, nullptr, false, nullptr, P.logits_ll, tb + hop);
    const uint32_t tag_logits = tb + hop;
    ++hop;
    // ---------------- sampler: CTA a < 8 handles attribute a (model.py:68-107); one numpy uniform per attribute per step
    if (c < 8) {
      const int a = c, o = meta.off[a], n = meta.off[a + 1] - meta.off[a];
      for (int i = tid; i < n; i += NCONS) {
        const float x = __uint_as_float(ll_wait(P.logits_ll + o + i, tag_logits, err));
        sm->sp[i] = x;
        if (P.logits_out) P.logits_out[o + i] = x;
      }
      cons_sync();
      const double u = P.uniforms[(size_t)t * 8 + a];
      const int tok = sample_attr(cx, n, meta.temp[a], meta.top_p[a], u);
      if (tid == 0) {
        P.sampled[(size_t)t * 8 + a] = tok;
        const int fed = P.forced ? P.forced[(size_t)t * 8 + a] : tok;
        ll_store(P.tok_ll + a, (uint32_t)fed, tb + hop);
      }
    }
    // ---------------- advance (model.py:59-65): every CTA applies the stop rule to the same 8 tokens
    if (tid < 8) sm->tok[tid] = (int)ll_wait(P.tok_ll + tid, tb + hop, err);
    cons_sync();
    bool stop = false;
#pragma unroll
    for (int a = 0; a < 8; ++a) { cur[a] = sm->tok[a]; stop |= cur[a] >= meta.pad[a]; }
    ++steps_run;
    if (!done) {
      if (stop) done = true;
      else if (c == 0 && tid < 8) P.result[(size_t)t * 8 + tid] = cur[tid];
    }
    if (c == 0 && tid == 0 && !done) P.n_written[0] = t + 1;
    cons_sync();
    if (done && P.stop_when_done) break;
  }
  // early exit: tell the producer to stop prefetching (it drains its in-flight copies)
  if (steps_run < n_steps && tid == 0) *reinterpret_cast<volatile int*>(&sm->stop) = 1;
  if (c == 0 && tid == 0) {
    *P.t_dev = t0 + steps_run;
    P.done[0] = done ? 1 : 0;
    *P.epoch = epoch0 + (uint32_t)max(n_steps, 1) * HOPS;
#pragma unroll
    for (int a = 0; a < 8; ++a) P.cur_tok[a] = cur[a];
  }
}

// cross-attention K/V of one layer from the projection layout [S_enc, 2 d] (K | V) to the split-major cache layout
__global__ void __launch_bounds__(128) decode_kv_relayout_kernel(const bf16* __restrict__ kv, bf16* __restrict__ k_out,
                                                                bf16* __restrict__ v_out, int S_enc) {
  pdl_entry();
  const int j = blockIdx.x, h = blockIdx.y;
  if (j >= S_enc) return;
  const size_t dst = ((size_t)(h * NSPLIT + j % NSPLIT) * NSLOT + j / NSPLIT) * HD + threadIdx.x;
  k_out[dst] = kv[(size_t)j * 2 * D + h * HD + threadIdx.x];
  v_out[dst] = kv[(size_t)j * 2 * D + D + h * HD + threadIdx.x];
}

}  // namespace pbdec

extern "C" int pb_decode_kv_relayout(const void* kv, void* k_out, void* v_out, int S_enc, void* stream) {
  using namespace pbdec;
  if (S_enc <= 0 || S_enc > NSPLIT * NSLOT) return pb_set_error("decode_kv_relayout: S_enc out of range");
  PB_LAUNCH(decode_kv_relayout_kernel, dim3(S_enc, H), 128, 0, reinterpret_cast<cudaStream_t>(stream), (const bf16*)kv,
            (bf16*)k_out, (bf16*)v_out, S_enc);
  return pb_check_launch("decode_kv_relayout");
}

extern "C" int pb_decode_persist_smem_bytes(void) { return (int)sizeof(pbdec::Shared) + 1024; }

extern "C" int pb_decode_persist_run(const pb_decode_persist_desc* d, int n_steps, const int* seg_sizes_host,
                                     const float* temp_host, const float* top_p_host, const int* pad_host, void* stream) {
  using namespace pbdec;
  if (d->n_layers < 1 || d->n_layers > MAXL) return pb_set_error("decode_persist: n_layers out of range");
  if (d->S_enc < 1 || d->S_enc > NSPLIT * NSLOT || d->S_max > NSPLIT * NSLOT) return pb_set_error("decode_persist: sequence too long");
  if (n_steps <= 0) return 0;
  SampleMeta m;
  int off = 0;
  for (int i = 0; i < 8; ++i) {
    m.off[i] = off; off += seg_sizes_host[i]; m.temp[i] = temp_host[i]; m.top_p[i] = top_p_host[i]; m.pad[i] = pad_host[i];
    if (seg_sizes_host[i] > 512) return pb_set_error("decode_persist: segment > 512");
  }
  m.off[8] = off;
  if (off != V) return pb_set_error("decode_persist: vocabulary must have 1280 entries");
  const int smem = (int)sizeof(Shared) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(decode_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return pb_set_cuda_error("cudaFuncSetAttribute(decode_persist)", e);
    attr = true;
  }
  int grid = pb_num_sms();
  if (grid < H * NSPLIT) return pb_set_error("decode_persist: needs at least 144 SMs");
  pb_decode_persist_desc dd = *d;
  void* args[] = {(void*)&dd, (void*)&m, (void*)&n_steps};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)decode_persist_kernel, dim3(grid), dim3(NTHREADS), args, (size_t)smem,
                                              reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return pb_set_cuda_error("cudaLaunchCooperativeKernel(decode_persist)", e);
  return pb_check_launch("decode_persist_kernel");
}

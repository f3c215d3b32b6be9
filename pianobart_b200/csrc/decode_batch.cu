// Persistent whole-step decode kernel for a BATCH of 64 sequences (BASELINE.json configs[2], batch 64; reference
// model.py:28-107 generalised to batches - the reference itself exits unless batch == 1), default model geometry.
//
// At batch 64 a decode step moves 3.4 GB through HBM, 95 % of it K/V cache (8 layers x 64 sequences x (self t + cross 1024)
// keys x 2 x 2 KB); the projections are 0.175 GB of weights and ~11 GFLOP.  Round 1 ran the step as a CUDA graph of 119
// kernels (split-K tcgen05 GEMMs with M = 64 + fp32 finalize kernels + split-key attention), 26 % of the HBM roofline:
// two thirds of the time went into the ~100 small kernels around the attention.  Here ONE cooperative launch runs N tokens:
//
//   * grid = one CTA per SM; 16 consumer warps + 1 producer warp per CTA; phases separated by a grid barrier (68 per token);
//   * projection phase: CTA c owns the output columns [8 u0, 8 u1) (u = c U / G, U = N / 8): the producer streams those weight
//     rows through a 4-slot shared-memory ring (bulk copies, padded rows -> conflict-free ldmatrix) ahead of time - weights
//     never wait for activations; the consumers stage the [64 x 1024] bf16 activation block from L2, apply the LayerNorm of
//     the previous sub-layer in shared memory, run mma.sync m16n8k16 (64 rows = 4 row tiles, K split over 4 warp groups) and
//     finish with bias / GELU / residual / position row in the epilogue - no split-K atomics, no finalize kernel;
//   * attention phase: the 512 (sequence, head) units are dealt round-robin to the CTAs; K then V of a unit stream through
//     the same ring in 64-key chunks (one contiguous 16 KB bulk copy each: the caches are laid out [seq][head][key][128]),
//     scores -> softmax -> P V in shared memory / registers, two passes, no partial results in HBM;
//   * sampler phase: the 512 (sequence, attribute) units, same arithmetic as the batch-1 kernel.
#include "decode_common.cuh"

namespace pbdec {
namespace bt {

constexpr int BM = 64;                       // sequences per launch = rows of every projection
constexpr int XSTRIDE = 1024 * 2 + 16;       // padded row stride (bytes): 8 consecutive rows start in 8 different 16-byte bank groups
constexpr int WSLOT = 8 * XSTRIDE;           // ring slot: 8 weight rows x 1024 columns (padded) or one 64-key K / V chunk (16 KB)
constexpr int BRING = 4;                     // weight ring (projection phases)
constexpr int KVRING = 4;                    // K / V ring of the attention phases: 4 x 32 KB slots in the (then idle) activation block
constexpr int KCH = 128;                     // keys per K / V chunk (one contiguous 32 KB bulk copy)
constexpr int KVSLOT = 2 * WSLOT;
constexpr int MAXKEYS = PB_DECODE_NSPLIT * PB_DECODE_NSLOT;   // 1026 >= 1024
constexpr int PART_STRIDE = 132;             // floats per partial attention result: max, sum, o[128] (+ pad)

struct BShared {
  alignas(1024) uint8_t x[BM * XSTRIDE];     // staged activations [64][1024] bf16 (padded rows); reduction scratch afterwards
  alignas(128) uint8_t ring[BRING][WSLOT];
  float sc[MAXKEYS + 62];                    // scores / probabilities of the unit in flight
  float q[HD], knew[HD], vnew[HD];
  float po[8][HD];
  float red[2 * NCW];
  float rowstat[BM][2];                      // per-row (sum, sum of squares) partials of an epilogue / (mean, rstd) of a staging pass
  float sp[512]; float sprob[512]; int sidx[512];
  int tok[8];
  alignas(8) uint64_t full_bar[BRING];
  alignas(8) uint64_t empty_bar[BRING];
  // The K / V chunks of an attention phase stream through the activation block x (unused then): 8 more slots, i.e. ~6
  // chunks = 96 KB in flight per SM (with the 4-slot ring alone the phase was latency bound at 29 GB/s per SM).  x_free: the
  // consumers tell the producer that the projection phase before an attention phase is done with x.
  alignas(8) uint64_t kv_full[KVRING];
  alignas(8) uint64_t kv_empty[KVRING];
  alignas(8) uint64_t x_free;
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 8-column units of CTA c for a projection with N output columns
__device__ __forceinline__ void unit_range(int N, int c, int G, int& u0, int& u1) {
  const unsigned U = (unsigned)N >> 3;
  u0 = (int)(((unsigned)c * U) / (unsigned)G);
  u1 = (int)(((unsigned)(c + 1) * U) / (unsigned)G);
}

struct BCtx {
  const pb_decode_batch_desc* p;
  BShared* sm;
  int c, G, tid, warp, lane;
  uint32_t chunk_no;
  uint32_t kv_no;        // running K / V chunk index (slot = kv_no % KVRING)
  uint32_t attn_no;      // attention phases entered so far (x_free phase)
  unsigned bar_gen;
  int* err;
  long long* tr;
  int tr_n;
  __device__ __forceinline__ const uint8_t* chunk_wait() {
    const uint32_t slot = chunk_no % BRING;
    mbar_wait_to(&sm->full_bar[slot], (chunk_no / BRING) & 1u, err);
    return sm->ring[slot];
  }
  __device__ __forceinline__ void chunk_release() {      // caller has synchronised the consumer threads
    if (tid == 0) pb::mbar_arrive(&sm->empty_bar[chunk_no % BRING]);
    ++chunk_no;
  }
  __device__ __forceinline__ const uint8_t* kv_wait() {
    const uint32_t slot = kv_no % KVRING;
    mbar_wait_to(&sm->kv_full[slot], (kv_no / KVRING) & 1u, err);
    return sm->x + (size_t)slot * KVSLOT;
  }
  __device__ __forceinline__ void kv_release() {
    if (tid == 0) pb::mbar_arrive(&sm->kv_empty[kv_no % KVRING]);
    ++kv_no;
  }
  // start of an attention phase (called after the grid barrier that ends the projection phase): x is free for K / V chunks
  __device__ __forceinline__ void attn_begin() {
    if (tid == 0) pb::mbar_arrive(&sm->x_free);
    ++attn_no;
  }
  // grid barrier over the consumer threads of all CTAs (the producer warps run ahead on their static schedule)
  __device__ __forceinline__ void grid_sync() {
    if (tr != nullptr && tid == 0 && tr_n < 160) tr[tr_n++] = clock64();
    cons_sync();
    bar_gen += (unsigned)G;
    if (tid == 0) {
      __threadfence();
      atomicAdd(p->barrier, 1u);
      if (ld_acquire_u32(p->barrier) < bar_gen) {
        const long long t0 = clock64();
        uint32_t n = 0;
        while (ld_acquire_u32(p->barrier) < bar_gen) {
          if ((++n & 1023u) == 0 && clock64() - t0 > TIMEOUT_CYCLES) die(err, 5);
        }
      }
    }
    cons_sync();
    if (tr != nullptr && tid == 0 && tr_n < 160) tr[tr_n++] = clock64();
  }
};

// ---------------------------------------------------------------------------------------------- projection phase
// X block: rows b = 0..63 of src (row stride ldx elements), columns [col0, col0 + 1024) -> shared memory (padded rows)
__device__ __forceinline__ void stage_x(BCtx& cx, const bf16* src, int ldx, int col0) {
  for (int i = cx.tid; i < BM * 128; i += NCONS) {        // 128 x 16 bytes per row
    const int r = i >> 7, q = i & 127;
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)r * ldx + col0) + q);
    *reinterpret_cast<uint4*>(cx.sm->x + (size_t)r * XSTRIDE + q * 16) = v;
  }
  cons_sync();
}
// Staging with the LayerNorm of the previous sub-layer applied on the fly: the row statistics (sum, sum of squares over
// the 1024 bf16-rounded values) were accumulated by the epilogues of the phase that produced `src` (fp32 atomics into
// `stats` [64][2]), so no CTA has to reduce the 64 rows again - the first version normalised all rows redundantly in every
// CTA and spent 13 k cycles per LayerNorm phase on it.  Row c of the result also goes to hn (residual of the sub-layer that
// follows): every CTA holds all rows, CTA c publishes row c.
__device__ __forceinline__ void stage_x_ln(BCtx& cx, const bf16* src, const float* stats, const float* gamma, const float* beta,
                                           bf16* hn) {
  BShared* sm = cx.sm;
  if (cx.tid < BM) {
    const float2 st = __ldcg(reinterpret_cast<const float2*>(stats) + cx.tid);
    const float mean = st.x * (1.0f / D);
    const float var = fmaxf(st.y * (1.0f / D) - mean * mean, 0.f);
    sm->rowstat[cx.tid][0] = mean;
    sm->rowstat[cx.tid][1] = rsqrtf(var + 1e-5f);
  }
  const int q = cx.tid & 127;                             // this thread's 8 columns (fixed over its 16 rows)
  float g[8], b[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + q * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + q * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + q * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + q * 8 + 4));
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  }
  uint4 raw[16];
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int r = (cx.tid >> 7) + 4 * it;
    raw[it] = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)r * D) + q);
  }
  cons_sync();                                            // row statistics visible
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int r = (cx.tid >> 7) + 4 * it;
    const float mean = sm->rowstat[r][0], rstd = sm->rowstat[r][1];
    const float2 f0 = unpack_bf16x2(raw[it].x), f1 = unpack_bf16x2(raw[it].y), f2 = unpack_bf16x2(raw[it].z), f3 = unpack_bf16x2(raw[it].w);
    uint4 w;
    w.x = pack_bf16x2((f0.x - mean) * rstd * g[0] + b[0], (f0.y - mean) * rstd * g[1] + b[1]);
    w.y = pack_bf16x2((f1.x - mean) * rstd * g[2] + b[2], (f1.y - mean) * rstd * g[3] + b[3]);
    w.z = pack_bf16x2((f2.x - mean) * rstd * g[4] + b[4], (f2.y - mean) * rstd * g[5] + b[5]);
    w.w = pack_bf16x2((f3.x - mean) * rstd * g[6] + b[6], (f3.y - mean) * rstd * g[7] + b[7]);
    *reinterpret_cast<uint4*>(sm->x + (size_t)r * XSTRIDE + q * 16) = w;
    if (r == cx.c) *(reinterpret_cast<uint4*>(hn + (size_t)r * D) + q) = w;
  }
  cons_sync();
}

// y[64, 8 u0 .. 8 u1) = epi( LN?(X)[64, K] W[8 u0 .. 8 u1, K]^T )      K = 1024 * npass
// epilogue: + bias, GELU, + resid[b, n] (bf16 [64, N]), + pos_row[n]; output bf16 [64, ldo] and / or fp32 [64, ldo]
// ln_stats != null: the staged block is LayerNorm(xsrc) (statistics from the producing phase); stats_out != null: this phase's
// epilogue accumulates the row statistics of ITS output for the LayerNorm that follows.
__device__ __forceinline__ void proj_phase(BCtx& cx, const bf16* xsrc, int ldx, int npass, int N, const float* ln_stats,
                                           const float* ln_g, const float* ln_b, bf16* hn_out, const float* bias, bool gelu,
                                           const bf16* resid, const bf16* pos_row, bf16* out_bf, float* out_f32, int ldo,
                                           float* stats_out) {
  BShared* sm = cx.sm;
  int u0, u1;
  unit_range(N, cx.c, cx.G, u0, u1);
  const int nu = u1 - u0;                                 // 0 .. 3
  const int mt = cx.warp & 3, kq = cx.warp >> 2;
  float acc[3][4];
#pragma unroll
  for (int u = 0; u < 3; ++u)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[u][i] = 0.f;
  const uint32_t xbase = pb::smem_u32(sm->x) + (uint32_t)((mt * 16 + (cx.lane & 15)) * XSTRIDE + (cx.lane >> 4) * 16);
  if (nu == 0 && (ln_stats == nullptr || cx.c >= BM)) return;   // nothing to compute, no row to publish (uniform over the CTA)
  for (int pass = 0; pass < npass; ++pass) {
    if (ln_stats != nullptr) stage_x_ln(cx, xsrc, ln_stats, ln_g, ln_b, hn_out);
    else stage_x(cx, xsrc, ldx, pass * 1024);
    uint32_t wb[3] = {0u, 0u, 0u};
    const uint32_t c0 = cx.chunk_no;
    for (int u = 0; u < nu; ++u) {
      const uint8_t* w = cx.chunk_wait();
      wb[u] = pb::smem_u32(w) + (uint32_t)((cx.lane & 7) * XSTRIDE + ((cx.lane >> 3) & 1) * 16);
      ++cx.chunk_no;
    }
    cx.chunk_no = c0;
    if (nu > 0) {
#pragma unroll 4
      for (int ks = 0; ks < 16; ++ks) {
        const uint32_t koff = (uint32_t)(kq * 256 + ks * 16) * 2u;
        uint32_t a0, a1, a2, a3;
        ldsm_x4(xbase + koff, a0, a1, a2, a3);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          if (u < nu) {
            uint32_t b0, b1;
            ldsm_x2(wb[u] + koff, b0, b1);
            mma_bf16(acc[u], a0, a1, a2, a3, b0, b1);
          }
        }
      }
    }
    cons_sync();                                          // every warp is done with X and the weight chunks of this pass
    for (int u = 0; u < nu; ++u) cx.chunk_release();
  }
  // cross-warp reduction over the 4 K quarters through shared memory (the X block is free now)
  float* scratch = reinterpret_cast<float*>(sm->x);      // [4][64][24]
  {
    const int g = cx.lane >> 2, t = cx.lane & 3;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      if (u < nu) {
        float* s0 = scratch + ((size_t)kq * BM + mt * 16 + g) * 24 + u * 8 + 2 * t;
        s0[0] = acc[u][0]; s0[1] = acc[u][1];
        s0[8 * 24] = acc[u][2]; s0[8 * 24 + 1] = acc[u][3];
      }
    }
  }
  cons_sync();
  const int ncp = nu * 4;                                 // column pairs of this CTA
  if (stats_out != nullptr && cx.tid < 2 * BM) (&sm->rowstat[0][0])[cx.tid] = 0.f;
  if (stats_out != nullptr) cons_sync();
  for (int i = cx.tid; i < BM * ncp; i += NCONS) {
    const int r = i / ncp, cp = i - r * ncp;
    const int col = 2 * cp, n0 = u0 * 8 + col;
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 pv = *reinterpret_cast<const float2*>(scratch + ((size_t)k * BM + r) * 24 + col);
      v0 += pv.x; v1 += pv.y;
    }
    if (bias) { v0 += __ldg(bias + n0); v1 += __ldg(bias + n0 + 1); }
    if (gelu) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
    if (resid) {
      const float2 rv = unpack_bf16x2(__ldcg(reinterpret_cast<const unsigned*>(resid + (size_t)r * N + n0)));
      v0 += rv.x; v1 += rv.y;
    }
    if (pos_row) { v0 += __bfloat162float(pos_row[n0]); v1 += __bfloat162float(pos_row[n0 + 1]); }
    if (out_bf) {
      const uint32_t w = pack_bf16x2(v0, v1);
      *reinterpret_cast<uint32_t*>(out_bf + (size_t)r * ldo + n0) = w;
      if (stats_out != nullptr) {                         // statistics of exactly what the next LayerNorm will read
        const float2 rb = unpack_bf16x2(w);
        atomicAdd(&sm->rowstat[r][0], rb.x + rb.y);
        atomicAdd(&sm->rowstat[r][1], rb.x * rb.x + rb.y * rb.y);
      }
    }
    if (out_f32) *reinterpret_cast<float2*>(out_f32 + (size_t)r * ldo + n0) = make_float2(v0, v1);
  }
  cons_sync();                                            // scratch is free before the next phase stages into it
  if (stats_out != nullptr && nu > 0 && cx.tid < 2 * BM) atomicAdd(stats_out + cx.tid, (&sm->rowstat[0][0])[cx.tid]);
}

// ---------------------------------------------------------------------------------------------- attention phase
// one (sequence b, head h) unit: nk cached keys streamed as K chunks then V chunks (+ the new key of this step for self
// attention); keep: encoder key-padding flags of the sequence or null.  Result -> out[b, h*128 .. +128) bf16.
// part >= 0: this call covers only a key range of the unit (the caller offsets `keep`; the producer streams that range): the
// unnormalised result (max, sum, o[128]) goes to P.part[part] and the second of the two halves to arrive (atomic counter, no
// waiting) merges both and writes `out`.
__device__ __forceinline__ void attn_unit(BCtx& cx, int nk, bool has_new, const uint8_t* keep, bf16* out, int part = -1) {
  BShared* sm = cx.sm;
  const int nkc = (nk + KCH - 1) / KCH;
  const float4 q4 = *reinterpret_cast<const float4*>(&sm->q[cx.lane * 4]);
  // scores: a warp takes 4 consecutive keys of a 64-key chunk at once (4 independent shuffle reductions in flight); two
  // 128-key chunks are processed per barrier (the first version did one key per warp at a time and one barrier per 64-key
  // chunk: 1.7 k cycles per 16 KB chunk = 18 GB/s per SM)
  for (int j = 0; j < nkc; j += 2) {
    const int nc = min(2, nkc - j);
    const uint8_t* kc[2];
    const uint32_t c0 = cx.kv_no;
    for (int u = 0; u < nc; ++u) { kc[u] = cx.kv_wait(); ++cx.kv_no; }
    cx.kv_no = c0;
    for (int u = 0; u < nc; ++u) {
      const int kbase = (j + u) * KCH, cnt = min(KCH, nk - kbase);
      for (int k0 = cx.warp * 4; k0 < cnt; k0 += NCW * 4) {
        float d[4];
        bool kp = true;
        if (keep != nullptr && cx.lane < 4 && k0 + cx.lane < cnt) kp = keep[kbase + k0 + cx.lane] != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint2 kw = *reinterpret_cast<const uint2*>(kc[u] + (size_t)min(k0 + i, cnt - 1) * (HD * 2) + cx.lane * 8);
          const float2 a = unpack_bf16x2(kw.x), b2 = unpack_bf16x2(kw.y);
          d[i] = q4.x * a.x + q4.y * a.y + q4.z * b2.x + q4.w * b2.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) d[i] += __shfl_xor_sync(0xffffffffu, d[i], o);
        }
        if (cx.lane < 4 && k0 + cx.lane < cnt) {
          const float dv = cx.lane == 0 ? d[0] : (cx.lane == 1 ? d[1] : (cx.lane == 2 ? d[2] : d[3]));
          sm->sc[kbase + k0 + cx.lane] = kp ? dv : -INFINITY;
        }
      }
    }
    cons_sync();
    for (int u = 0; u < nc; ++u) cx.kv_release();
  }
  if (has_new && cx.warp == 0) {
    const float4 k4 = *reinterpret_cast<const float4*>(&sm->knew[cx.lane * 4]);
    float d = q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (cx.lane == 0) sm->sc[nk] = d;
  }
  cons_sync();
  const int n = nk + (has_new ? 1 : 0);
  float m = -INFINITY;
  for (int i = cx.tid; i < n; i += NCONS) m = fmaxf(m, sm->sc[i]);
  m = cons_max(m, sm->red);
  float l = 0.f;
  for (int i = cx.tid; i < n; i += NCONS) {
    const float e = (m == -INFINITY) ? 0.f : __expf(sm->sc[i] - m);
    sm->sc[i] = e;
    l += e;
  }
  l = cons_sum(l, sm->red);                               // (contains the barrier that publishes the probabilities)
  const int g = cx.tid >> 6, dp = cx.tid & 63;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int j = 0; j < nkc; j += 2) {
    const int nc = min(2, nkc - j);
    const uint8_t* vc[2];
    const uint32_t c0 = cx.kv_no;
    for (int u = 0; u < nc; ++u) { vc[u] = cx.kv_wait(); ++cx.kv_no; }
    cx.kv_no = c0;
    for (int u = 0; u < nc; ++u) {
      const int kbase = (j + u) * KCH, cnt = min(KCH, nk - kbase);
#pragma unroll 4
      for (int k = g; k < cnt; k += 16) {
        const float2 v2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vc[u] + (size_t)k * (HD * 2) + dp * 4));
        const float pj = sm->sc[kbase + k];
        a0 += pj * v2.x;
        a1 += pj * v2.y;
        if (k + 8 < cnt) {
          const float2 w2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(vc[u] + (size_t)(k + 8) * (HD * 2) + dp * 4));
          const float pk = sm->sc[kbase + k + 8];
          a2 += pk * w2.x;
          a3 += pk * w2.y;
        }
      }
    }
    cons_sync();
    for (int u = 0; u < nc; ++u) cx.kv_release();
  }
  a0 += a2; a1 += a3;
  sm->po[g][2 * dp] = a0;
  sm->po[g][2 * dp + 1] = a1;
  cons_sync();
  if (cx.tid < HD / 2) {
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) { o0 += sm->po[gg][2 * cx.tid]; o1 += sm->po[gg][2 * cx.tid + 1]; }
    if (has_new) { o0 += sm->sc[nk] * sm->vnew[2 * cx.tid]; o1 += sm->sc[nk] * sm->vnew[2 * cx.tid + 1]; }
    if (part < 0) {
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      *reinterpret_cast<uint32_t*>(out + 2 * cx.tid) = pack_bf16x2(o0 * inv, o1 * inv);
    } else {
      float* gp = cx.p->part + (size_t)part * PART_STRIDE;
      if (cx.tid == 0) { gp[0] = m; gp[1] = l; }
      gp[2 + 2 * cx.tid] = o0;
      gp[3 + 2 * cx.tid] = o1;
      __threadfence();
    }
  }
  cons_sync();
  if (part >= 0) {
    if (cx.tid == 0) sm->tok[0] = atomicAdd(cx.p->part_cnt + (part >> 1), 1);
    cons_sync();
    const int old = sm->tok[0];
    cons_sync();
    if (old == 1) {                                        // both halves are in global memory: merge them
      __threadfence();
      if (cx.tid < HD / 2) {
        const float* g0 = cx.p->part + (size_t)(part & ~1) * PART_STRIDE;
        const float* g1 = g0 + PART_STRIDE;
        const float m0 = __ldcg(g0), l0 = __ldcg(g0 + 1), m1 = __ldcg(g1), l1 = __ldcg(g1 + 1);
        const float mm = fmaxf(m0, m1);
        const float w0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - mm), w1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - mm);
        const float ll = l0 * w0 + l1 * w1;
        const float inv = ll > 0.f ? 1.0f / ll : 0.f;
        const float o0 = __ldcg(g0 + 2 + 2 * cx.tid) * w0 + __ldcg(g1 + 2 + 2 * cx.tid) * w1;
        const float o1 = __ldcg(g0 + 3 + 2 * cx.tid) * w0 + __ldcg(g1 + 3 + 2 * cx.tid) * w1;
        *reinterpret_cast<uint32_t*>(out + 2 * cx.tid) = pack_bf16x2(o0 * inv, o1 * inv);
      }
      if (cx.tid == 0) cx.p->part_cnt[part >> 1] = 0;      // next use: the next cross-attention phase, grid barriers away
    }
  }
}

// Work list of an attention phase: CTA c takes units c, c + G, ... while whole rounds last; the units of the last, partial
// round are cut into two key halves (chunk-aligned) when that gives every CTA at most one piece - 512 units on 148 CTAs:
// 3 whole units + one half instead of 4 units for 68 CTAs and 3 for the rest (a quarter of the phase was idle time).
struct AttnItem { int ui, k_lo, k_n, part; };
__device__ __forceinline__ int attn_items(int c, int G, int nunits, int nk, bool can_split, AttnItem (&it)[8]) {
  const int nfull = nunits / G, rem = nunits - nfull * G;
  const int nkc = (nk + KCH - 1) / KCH;
  int n = 0;
  for (int i = 0; i < nfull && n < 7; ++i) it[n++] = AttnItem{c + i * G, 0, nk, -1};
  if (can_split && rem > 0 && 2 * rem <= G && nkc >= 2) {
    if (c < 2 * rem) {
      const int h0 = ((nkc + 1) / 2) * KCH, half = c & 1;
      it[n++] = AttnItem{nfull * G + (c >> 1), half ? h0 : 0, half ? nk - h0 : h0, c};
    }
  } else if (c < rem) {
    it[n++] = AttnItem{nfull * G + c, 0, nk, -1};
  }
  return n;
}

// ---------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(NTHREADS, 1) decode_batch_kernel(const __grid_constant__ pb_decode_batch_desc P,
                                                                    const SampleMeta meta, int n_steps) {
  extern __shared__ uint8_t smem_raw[];
  BShared* sm = reinterpret_cast<BShared*>(smem_raw + ((1024u - (pb::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.x, G = gridDim.x;
  const int nl = P.n_layers, Se = P.S_enc, Smax = P.S_max;
  int* err = P.error_flag;
  if (tid == 0) {
    for (int i = 0; i < BRING; ++i) { pb::mbar_init(&sm->full_bar[i], 1); pb::mbar_init(&sm->empty_bar[i], 1); }
    for (int i = 0; i < KVRING; ++i) { pb::mbar_init(&sm->kv_full[i], 1); pb::mbar_init(&sm->kv_empty[i], 1); }
    pb::mbar_init(&sm->x_free, 1);
    pb::fence_mbar_init();
  }
  __syncthreads();
  const int t0 = *P.t_dev;
  n_steps = max(0, min(n_steps, Smax - t0));
  const int nunits = BM * H;                               // (sequence, head) attention units; unit i -> b = i / 8, h = i % 8
  const bool split_ok = P.part != nullptr && P.part_cnt != nullptr;   // scratch for the halves of the last, partial round

  if (warp == NCW) {
    // ============================================================ producer: weights and K / V chunks, static schedule
    if (lane == 0) {
      const uint64_t pol_w = l2_policy_evict_first(), pol_kv = l2_policy_evict_first();
      uint32_t k = 0;
      auto slot_wait = [&]() {
        const uint32_t slot = k % BRING;
        if (k >= BRING) mbar_wait_to(&sm->empty_bar[slot], ((k / BRING) & 1u) ^ 1u, err);
        return slot;
      };
      auto proj = [&](const void* W, int N, int npass) {    // K = 1024 * npass
        int u0, u1;
        unit_range(N, c, G, u0, u1);
        for (int pass = 0; pass < npass; ++pass)
          for (int u = u0; u < u1; ++u) {
            const uint32_t slot = slot_wait();
            pb::mbar_expect_tx(&sm->full_bar[slot], 8 * 2048);
            const bf16* src = reinterpret_cast<const bf16*>(W) + (size_t)(8 * u) * (1024 * npass) + pass * 1024;
#pragma unroll
            for (int r = 0; r < 8; ++r)
              bulk_load(sm->ring[slot] + r * XSTRIDE, src + (size_t)r * (1024 * npass), 2048, &sm->full_bar[slot], pol_w);
            ++k;
          }
      };
      // attention streams K (all chunks) then V (all chunks) per unit: interleave per unit
      uint32_t kk = 0, na = 0;
      auto attn = [&](const void* kb, const void* vb, int row_cap, int nk_all, bool can_split) {
        // the consumers have left the projection phase that used x (their arrival follows its closing grid barrier)
        mbar_wait_to(&sm->x_free, na & 1u, err);
        ++na;
        AttnItem items[8];
        const int nit = attn_items(c, G, nunits, nk_all, can_split, items);
        for (int ii = 0; ii < nit; ++ii) {
          const int ui = items[ii].ui, nk = items[ii].k_n;
          const int nkc = (nk + KCH - 1) / KCH;
          for (int pass = 0; pass < 2; ++pass) {
            const bf16* ub = reinterpret_cast<const bf16*>(pass == 0 ? kb : vb) + ((size_t)ui * row_cap + items[ii].k_lo) * HD;
            for (int j = 0; j < nkc; ++j) {
              const uint32_t slot = kk % KVRING;
              if (kk >= KVRING) mbar_wait_to(&sm->kv_empty[slot], ((kk / KVRING) & 1u) ^ 1u, err);
              const uint32_t bytes = (uint32_t)min(KCH, nk - j * KCH) * HD * 2;
              pb::mbar_expect_tx(&sm->kv_full[slot], bytes);
              bulk_load(sm->x + (size_t)slot * KVSLOT, ub + (size_t)j * KCH * HD, bytes, &sm->kv_full[slot], pol_kv);
              ++kk;
            }
          }
        }
      };
      for (int st = 0; st < n_steps; ++st) {
        const int t = t0 + st;
        proj(P.w_in, D, 2);
        for (int l = 0; l < nl; ++l) {
          const pb_decode_layer& L = P.layer[l];
          proj(L.wqkv, 3 * D, 1);
          attn(L.self_k, L.self_v, Smax, t, false);
          proj(L.wo, D, 1);
          proj(L.wqc, D, 1);
          attn(L.cross_k, L.cross_v, Se, Se, split_ok);
          proj(L.woc, D, 1);
          proj(L.w1, F, 1);
          proj(L.w2, D, 2);
        }
        proj(P.w_heads, V, 1);
      }
    }
    return;
  }

  // ============================================================== consumers
  BCtx cx;
  cx.p = &P; cx.sm = sm; cx.c = c; cx.G = G; cx.tid = tid; cx.warp = warp; cx.lane = lane;
  cx.chunk_no = 0; cx.kv_no = 0; cx.attn_no = 0; cx.bar_gen = 0; cx.err = err; cx.tr = nullptr; cx.tr_n = 0;
  bf16* xemb = reinterpret_cast<bf16*>(P.xemb);
  bf16* raw0 = reinterpret_cast<bf16*>(P.raw0); bf16* raw1 = reinterpret_cast<bf16*>(P.raw1); bf16* raw2 = reinterpret_cast<bf16*>(P.raw2);
  bf16* hn = reinterpret_cast<bf16*>(P.hn); bf16* qkvb = reinterpret_cast<bf16*>(P.qkv); bf16* qcb = reinterpret_cast<bf16*>(P.qc);
  bf16* ob = reinterpret_cast<bf16*>(P.ob); bf16* f1 = reinterpret_cast<bf16*>(P.f1);
  float* st0 = P.stats; float* st1 = P.stats + 2 * BM; float* st2 = P.stats + 4 * BM;   // row statistics of raw0 / raw1 / raw2
  const float qscale = 0.08838834764831845f;              // hd^-0.5

  for (int st = 0; st < n_steps; ++st) {
    const int t = t0 + st;
    cx.tr = (P.trace != nullptr && st == n_steps - 1) ? P.trace + (size_t)c * 160 : nullptr;
    // ---------------- advance of the PREVIOUS step + embedding of its token (cur_tok was written by the sampler phase)
    if (c < BM) {
      // row c: 8 embedding rows (table pre-scaled by 16) -> xemb[c, 2048]
      for (int e2 = tid; e2 < E / 2; e2 += NCONS) {
        const int a = e2 >> 7, col = (e2 & 127) * 2;
        const int tok = __ldcg(P.cur_tok + c * 8 + a);
        const uint32_t w = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const bf16*>(P.emb_table) +
                                                              (size_t)(meta.off[a] + tok) * 256 + col);
        *reinterpret_cast<uint32_t*>(xemb + (size_t)c * E + 2 * e2) = w;
      }
    }
    cx.grid_sync();
    proj_phase(cx, xemb, E, 2, D, nullptr, nullptr, nullptr, nullptr, P.b_in, false, nullptr,
               reinterpret_cast<const bf16*>(P.pos_table) + (size_t)(t + 2) * D, raw0, nullptr, D, st0);
    cx.grid_sync();
    const float* ln_g = P.lne_g; const float* ln_b = P.lne_b;
    for (int l = 0; l < nl; ++l) {
      const pb_decode_layer& L = P.layer[l];
      // ---- qkv = LN(raw0) Wqkv^T + b
      proj_phase(cx, raw0, D, 1, 3 * D, st0, ln_g, ln_b, hn, L.bqkv, false, nullptr, nullptr, qkvb, nullptr, 3 * D, nullptr);
      cx.grid_sync();
      if (c == 0 && tid < 2 * BM) st0[tid] = 0.f;         // consumed by every CTA; accumulated again by fc2 (6 barriers from here)
      // ---- self attention (append key t, attend keys 0..t)
      cx.attn_begin();
      for (int ui = c; ui < nunits; ui += G) {
        const int b = ui >> 3, h = ui & 7;
        if (tid < 3 * (HD / 2)) {
          const int which = tid / (HD / 2), w = tid % (HD / 2);
          const uint32_t word = __ldcg(reinterpret_cast<const unsigned*>(qkvb + (size_t)b * 3 * D + which * D + h * HD) + w);
          const float2 f = unpack_bf16x2(word);
          float* dst = which == 0 ? sm->q : (which == 1 ? sm->knew : sm->vnew);
          const float scl = which == 0 ? qscale : 1.0f;
          dst[2 * w] = f.x * scl;
          dst[2 * w + 1] = f.y * scl;
          if (which > 0) {
            bf16* dstc = reinterpret_cast<bf16*>(which == 1 ? L.self_k : L.self_v) + ((size_t)ui * Smax + t) * HD;
            *reinterpret_cast<uint32_t*>(dstc + 2 * w) = word;
          }
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");      // later bulk copies (async proxy) read the appended rows
        cons_sync();
        attn_unit(cx, t, true, nullptr, ob + (size_t)b * D + h * HD);
      }
      cx.grid_sync();
      // ---- raw1 = ob Wo^T + bo + h
      proj_phase(cx, ob, D, 1, D, nullptr, nullptr, nullptr, nullptr, L.bo, false, hn, nullptr, raw1, nullptr, D, st1);
      cx.grid_sync();
      // ---- q_c = LN1(raw1) Wqc^T + b
      proj_phase(cx, raw1, D, 1, D, st1, L.ln1_g, L.ln1_b, hn, L.bqc, false, nullptr, nullptr, qcb, nullptr, D, nullptr);
      cx.grid_sync();
      if (c == 0 && tid < 2 * BM) st1[tid] = 0.f;
      // ---- cross attention over the encoder keys
      cx.attn_begin();
      {
        AttnItem items[8];
        const int nit = attn_items(c, G, nunits, Se, split_ok, items);
        for (int ii = 0; ii < nit; ++ii) {
          const int ui = items[ii].ui, b = ui >> 3, h = ui & 7;
          if (tid < HD / 2) {
            const float2 f = unpack_bf16x2(__ldcg(reinterpret_cast<const unsigned*>(qcb + (size_t)b * D + h * HD) + tid));
            sm->q[2 * tid] = f.x * qscale;
            sm->q[2 * tid + 1] = f.y * qscale;
          }
          cons_sync();
          attn_unit(cx, items[ii].k_n, false, P.enc_keep ? P.enc_keep + (size_t)b * Se + items[ii].k_lo : nullptr,
                    ob + (size_t)b * D + h * HD, items[ii].part);
        }
      }
      cx.grid_sync();
      proj_phase(cx, ob, D, 1, D, nullptr, nullptr, nullptr, nullptr, L.boc, false, hn, nullptr, raw2, nullptr, D, st2);
      cx.grid_sync();
      proj_phase(cx, raw2, D, 1, F, st2, L.ln2_g, L.ln2_b, hn, L.b1, true, nullptr, nullptr, f1, nullptr, F, nullptr);
      cx.grid_sync();
      if (c == 0 && tid < 2 * BM) st2[tid] = 0.f;
      proj_phase(cx, f1, F, 2, D, nullptr, nullptr, nullptr, nullptr, L.b2, false, hn, nullptr, raw0, nullptr, D, st0);
      cx.grid_sync();
      ln_g = L.ln3_g; ln_b = L.ln3_b;
    }
    // ---------------- heads -> fp32 logits
    proj_phase(cx, raw0, D, 1, V, st0, ln_g, ln_b, hn, P.b_heads, false, nullptr, nullptr, nullptr, P.logits, V, nullptr);
    cx.grid_sync();
    if (c == 0 && tid < 2 * BM) st0[tid] = 0.f;
    // ---------------- sampler: (sequence, attribute) units
    for (int ui = c; ui < BM * 8; ui += G) {
      const int b = ui >> 3, a = ui & 7;
      const int o = meta.off[a], n = meta.off[a + 1] - meta.off[a];
      for (int i = tid; i < n; i += NCONS) sm->sp[i] = __ldcg(P.logits + (size_t)b * V + o + i);
      cons_sync();
      const double u = P.uniforms[((size_t)b * Smax + t) * 8 + a];
      const int tok = sample_core(tid, n, meta.temp[a], meta.top_p[a], u, sm->sp, sm->sprob, sm->sidx, sm->red, &sm->tok[0]);
      if (tid == 0) {
        P.sampled[((size_t)b * Smax + t) * 8 + a] = tok;
        P.cur_tok[b * 8 + a] = P.forced ? P.forced[((size_t)b * Smax + t) * 8 + a] : tok;
      }
      cons_sync();
    }
    cx.grid_sync();
    // ---------------- advance (model.py:59-65), one thread per sequence
    if (c == 0 && tid < BM) {
      const int b = tid;
      if (!P.done[b]) {
        bool stop = false;
        int tk[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) { tk[a] = __ldcg(P.cur_tok + b * 8 + a); stop |= tk[a] >= meta.pad[a]; }
        if (stop) P.done[b] = 1;
        else {
#pragma unroll
          for (int a = 0; a < 8; ++a) P.result[((size_t)b * Smax + t) * 8 + a] = tk[a];
          P.n_written[b] = t + 1;
        }
      }
    }
  }
  if (c == 0 && tid == 0) *P.t_dev = t0 + n_steps;
}

// cross-attention K / V of one layer: projection layout [B * S_enc, 2 d] (K | V) -> [B][8][S_enc][128] each
__global__ void __launch_bounds__(128) decode_kv_relayout_batch_kernel(const bf16* __restrict__ kv, bf16* __restrict__ k_out,
                                                                      bf16* __restrict__ v_out, int S_enc) {
  pdl_entry();
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const size_t src = ((size_t)b * S_enc + j) * 2 * D + h * HD + threadIdx.x;
  const size_t dst = (((size_t)b * H + h) * S_enc + j) * HD + threadIdx.x;
  k_out[dst] = kv[src];
  v_out[dst] = kv[src + D];
}

}  // namespace bt
}  // namespace pbdec

extern "C" int pb_decode_kv_relayout_batch(const void* kv, void* k_out, void* v_out, int B, int S_enc, void* stream) {
  using namespace pbdec;
  if (B <= 0 || S_enc <= 0) return pb_set_error("decode_kv_relayout_batch: empty");
  PB_LAUNCH(bt::decode_kv_relayout_batch_kernel, dim3(S_enc, H, B), 128, 0, reinterpret_cast<cudaStream_t>(stream),
            (const bf16*)kv, (bf16*)k_out, (bf16*)v_out, S_enc);
  return pb_check_launch("decode_kv_relayout_batch");
}

extern "C" int pb_decode_batch_run(const pb_decode_batch_desc* d, int n_steps, const int* seg_sizes_host, const float* temp_host,
                                   const float* top_p_host, const int* pad_host, void* stream) {
  using namespace pbdec;
  if (d->n_layers < 1 || d->n_layers > MAXL) return pb_set_error("decode_batch: n_layers out of range");
  if (d->B != bt::BM) return pb_set_error("decode_batch: batch must be 64");
  if (d->S_enc < 1 || d->S_enc > bt::MAXKEYS || d->S_max > bt::MAXKEYS) return pb_set_error("decode_batch: sequence too long");
  if (n_steps <= 0) return 0;
  SampleMeta m;
  int off = 0;
  for (int i = 0; i < 8; ++i) {
    m.off[i] = off; off += seg_sizes_host[i]; m.temp[i] = temp_host[i]; m.top_p[i] = top_p_host[i]; m.pad[i] = pad_host[i];
    if (seg_sizes_host[i] > 512) return pb_set_error("decode_batch: segment > 512");
  }
  m.off[8] = off;
  if (off != V) return pb_set_error("decode_batch: vocabulary must have 1280 entries");
  const int smem = (int)sizeof(bt::BShared) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(bt::decode_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return pb_set_cuda_error("cudaFuncSetAttribute(decode_batch)", e);
    attr = true;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(d->barrier, 0, sizeof(unsigned), st);   // the grid-barrier counter starts at 0 in every launch
  if (e != cudaSuccess) return pb_set_cuda_error("cudaMemsetAsync(decode_batch barrier)", e);
  const int grid = pb_num_sms();
  pb_decode_batch_desc dd = *d;
  void* args[] = {(void*)&dd, (void*)&m, (void*)&n_steps};
  e = cudaLaunchCooperativeKernel((const void*)bt::decode_batch_kernel, dim3(grid), dim3(NTHREADS), args, (size_t)smem, st);
  if (e != cudaSuccess) return pb_set_cuda_error("cudaLaunchCooperativeKernel(decode_batch)", e);
  return pb_check_launch("decode_batch_kernel");
}

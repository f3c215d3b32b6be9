// Fused (flash-style) attention for sm_100a: tcgen05 MMAs with S / dP / accumulators in TMEM, operands staged
// by TMA, masks derived in-kernel (key-padding bitmap + causal index compare) - no (B,H,S,S) tensor ever
// reaches HBM.  Replaces HF BartAttention's core (eager_attention_forward: softmax(QK^T*hd^-0.5 + mask) V) and
// its autograd for the three PianoBART variants: encoder self (key padding), decoder self (causal & padding),
// decoder cross (encoder key padding).  head_dim is fixed at 128 (default model).
//
// Three kernels share one tile convention: every operand tile is [128 rows x 128 cols] bf16 stored as two
// 64-column halves of [128 x 128 B] with the 128-byte swizzle.  The same physical tile can be consumed
//   * K-major  : rows = M/N index, columns = contraction index  (Q, K for S=QK^T; dO, V for dP=dO V^T)
//   * MN-major : rows = contraction index, columns = M/N index  (V for O=PV; K for dQ=dS K; dO, Q for dV, dK)
// P, dS (and their transposes) never touch shared memory: the softmax warps write them back (bf16, packed) over the
// TMEM columns of the S / dP values they were derived from, and the next MMA reads them as its A operand from TMEM.
//
//   attn_fwd3    CTA = (b, h, 2 x 128 queries) loop over 128-key blocks: S=QK^T -> online softmax -> O += P V ; LSE out
//                (the production forward kernel; attn_fwd is its one-tile predecessor, PIANOBART_B200_ATTN_FWD=1)
//   attn_bwd_dkv CTA = (b, h, 128 keys)     loop over 64-query blocks: S^T, dP^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q
//   attn_bwd_dq  CTA = (b, h, 128 queries)  loop over 128-key blocks:  S, dP -> dS -> dQ += dS K
//   attn_bwd_prep                           D = rowsum(dO * O)
// Warp roles per CTA of the backward kernels and attn_fwd (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9
// softmax / epilogue: thread <-> (TMEM lane = tile row, 64-column half), two warps per scheduler.
// Measured facts the pipelines are built around (tools/micro/mma_bench.cu, -DPB_TRACE phase traces): one tcgen05.mma
// (M=128, K=16) holds the tensor pipe for max(~72, N/2) cycles whatever the operand source; a TMA tile refill takes
// 1500-2300 cycles under load; the MUFU pipe (ex2) is what bounds the softmax.  Outputs leave through bulk tensor
// stores from an operand tile that is free by then.
#include "ptx.cuh"
#include "pb_internal.h"
#include <stdlib.h>
#include <type_traits>

namespace pb {

constexpr int AT = 128;                 // tile edge (queries, keys and head_dim)
constexpr int TILE_BYTES = AT * AT * 2;  // 32 KB
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr float LOG2E = 1.4426950408889634f;

// Developer-only phase trace (build with PIANOBART_B200_NVCC_EXTRA=-DPB_TRACE): clock64 stamps of one CTA's MMA warp
// and two softmax warps per block, read back with pb_debug_trace().
#ifdef PB_TRACE
__device__ long long pb_trace_buf[3 * 64 * 8];
#define PB_TRACE_ON (blockIdx.x == 3 && blockIdx.y == 3 && blockIdx.z == 8)
#define PB_TR(role, j, k) do { if (PB_TRACE_ON && (j) < 64) pb_trace_buf[((role) * 64 + (j)) * 8 + (k)] = clock64(); } while (0)
#else
#define PB_TR(role, j, k) do { } while (0)
#endif

struct AttnParams {
  int B, H, Sq, Sk;
  int causal;
  float scale;                      // hd^-0.5
  const uint8_t* key_keep;          // [B, Sk] or null
  // forward
  __nv_bfloat16* o; long long ldo;  // [B, Sq, H*hd]
  float* lse;                       // [B, H, Sq]  (log2 domain: max + log2(sum))
  // backward
  const float* dvec;                // D [B, H, Sq]
  __nv_bfloat16* dq; long long lddq;
  __nv_bfloat16* dk; long long lddk;
  __nv_bfloat16* dv; long long lddv;
  long long dq_sb, dk_sb, dv_sb, o_sb;  // batch strides (elements)
  int dbg_delay;                    // test hook (pb_debug_set_attn_delay): TMA producers stall up to this many cycles per load
};

// Fault injection for the pipeline tests: the producer thread spins a pseudo-random number of cycles (< max_cycles) before a
// load, so that tiles arrive late / out of their usual order relative to the softmax warps and the MMA thread.
__device__ __forceinline__ void chaos_delay(int max_cycles, uint32_t salt) {
  if (max_cycles <= 0) return;
  uint32_t h = salt * 2654435761u + blockIdx.x * 40503u + blockIdx.y * 9973u + blockIdx.z * 101u;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  const long long until = clock64() + (long long)(h % (uint32_t)max_cycles);
  while (clock64() < until) { }
}

// ---- tile helpers --------------------------------------------------------------------------------------
// K-major view of a tile: k-step kk (16 contraction columns) of 8
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int kk) {
  return make_smem_desc_sw128(tile + (kk >> 2) * HALF_BYTES + (kk & 3) * 32, 16, 1024);
}
// MN-major view: contraction index = tile row; k-step kk = 16 rows = 2048 B; the two 64-wide MN halves are
// HALF_BYTES apart (LBO), 8-row groups 1024 B apart (SBO)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int kk) {
  return make_smem_desc_sw128(tile + kk * 2048, HALF_BYTES, 1024);
}
// D[tmem] (+)= A * B over the full 128-deep contraction
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void mma_tile(uint32_t tmem_d, uint32_t a_tile, uint32_t b_tile, bool accumulate) {
  constexpr uint32_t idesc = make_idesc_bf16(AT, AT, A_MN ? 1 : 0, B_MN ? 1 : 0);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint64_t ad = A_MN ? desc_mnmajor(a_tile, kk) : desc_kmajor(a_tile, kk);
    const uint64_t bd = B_MN ? desc_mnmajor(b_tile, kk) : desc_kmajor(b_tile, kk);
    umma_bf16(tmem_d, ad, bd, idesc, (accumulate || kk > 0) ? 1u : 0u);
  }
}
// TMA load of one [128 x 128] tile (two 64-column boxes) at (row0, h, b)
__device__ __forceinline__ void load_tile(uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int row0, int h, int b) {
  tma_load_4d(dst, m, bar, 0, row0, h, b);
  tma_load_4d(dst + HALF_BYTES, m, bar, 64, row0, h, b);
}
// store 32 consecutive columns [c0, c0+32) of row r (bf16) into a swizzled tile
__device__ __forceinline__ void store_chunk(uint8_t* tile, int r, int c0, const float (&x)[32]) {
  uint8_t* half = tile + (c0 >> 6) * HALF_BYTES + r * 128;
  const int cbase = (c0 & 63) >> 3;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 v;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(x[g * 8 + 2 * t], x[g * 8 + 2 * t + 1]);
    *reinterpret_cast<uint4*>(half + (((cbase + g) ^ (r & 7)) << 4)) = v;
  }
}
// Output epilogue: accumulator rows are staged (bf16) in a shared-memory tile that is free by then, in the same two-half
// 128-byte-swizzled layout the operand tiles use, and written with two bulk tensor stores - a thread owns one row, so direct
// global stores would cost 32 L1 wavefronts per warp instruction (3200 cycles per CTA in the forward kernel, profiles/).
// Rows beyond the sequence are clipped by the tensor map.
__device__ __forceinline__ void stage_row_chunk(uint8_t* tile, int r, int hf, int c, const uint32_t (&v)[32], float scale) {
  uint8_t* rowp = tile + hf * HALF_BYTES + r * 128;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 o;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int tt = 0; tt < 4; ++tt)
      h2[tt] = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * tt]) * scale, __uint_as_float(v[g * 8 + 2 * tt + 1]) * scale);
    *reinterpret_cast<uint4*>(rowp + (((c * 4 + g) ^ (r & 7)) << 4)) = o;
  }
}
__device__ __forceinline__ void store_tile_tma(const CUtensorMap* m, const uint8_t* tile, int row0, int h, int b) {
  tma_store_4d(m, tile, 0, row0, h, b);
  tma_store_4d(m, tile + HALF_BYTES, 64, row0, h, b);
  bulk_commit();
}

constexpr int NCOMPUTE = 256;             // 8 softmax / epilogue warps: (TMEM lane quadrant) x (column half)
constexpr int NTHREADS = 64 + NCOMPUTE;
__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 pairs (Blackwell FFMA2 / FADD2): one issue slot for two lanes of the softmax's scale-and-shift and row sums
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// 2^y for a pair of log2-domain arguments on the FMA / integer pipes (no MUFU): round-to-nearest split y = k + f with the
// magic-number trick, degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16
// rounding of P), k added into the exponent field.  Arguments below -120 give 2^-120 (0 after the bf16 P V product).  The
// softmax sends a share of the pairs of a row here so that the MUFU pipe (ex2: ~10.5 cycles per warp instruction measured on
// B200, i.e. ~1340 cycles per 128 x 128 block) and the FMA pipe work on the exponentials of a block side by side.  With one
// softmax warp per scheduler the gain saturates early (a polynomial pair costs 10 extra issue slots at ~0.5 IPC, about what
// its two MUFU slots cost): 1450 cycles per block without it, 1250 at 25-37 %, 1420 at 50 % (profiles/r2_summary.md).
__device__ __forceinline__ float2 exp2_poly2(float2 y) {
  y.x = fmaxf(y.x, -120.f);
  y.y = fmaxf(y.y, -120.f);
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
  const float2 t = fadd2(y, magic);
  const float2 r = fadd2(t, nmagic);
  const float2 f = fadd2(y, make_float2(-r.x, -r.y));
  float2 p = ffma2(make_float2(0.0551716685295105f, 0.0551716685295105f), f, make_float2(0.2426111251115799f, 0.2426111251115799f));
  p = ffma2(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = ffma2(p, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
  float2 e;
  e.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  e.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return e;
}
// Backward softmax of an unmasked pair: P = 2^(S sl2 + negL), dS = P (dP scale + negDs), in packed fp32; POLY sends the two
// exponentials to the FMA pipe (exp2_poly2) instead of MUFU.
#ifndef PB_BWD_POLY_MASK
#define PB_BWD_POLY_MASK 0        // measured (profiles/r2_attn_bwd_poly.log, B16 H8 S1024): scalar 0.284 ms, packed 0.278,
                                  // packed + 5/16 polynomial 0.279, + 8/16 0.283 - the backward kernels are not MUFU-bound
#endif
template <bool POLY>
__device__ __forceinline__ void bwd_pair(uint32_t s0, uint32_t s1, uint32_t d0, uint32_t d1, float2 sl2v, float2 negL,
                                         float2 scv, float2 negDs, float2& e, float2& ds) {
  const float2 y = ffma2(make_float2(__uint_as_float(s0), __uint_as_float(s1)), sl2v, negL);
  if (POLY) { e = exp2_poly2(y); } else { e.x = ex2(y.x); e.y = ex2(y.y); }
  ds = fmul2(e, ffma2(make_float2(__uint_as_float(d0), __uint_as_float(d1)), scv, negDs));
}
// bit i of the result = column (c0 + i) of this key block may be attended by query qg:
// key-padding bitmap word AND (causal: kg0 + c0 + i <= qg)
__device__ __forceinline__ uint32_t chunk_mask(uint32_t keep_word, bool causal, int qg, int kg_c0) {
  if (!causal) return keep_word;
  const int lim = qg - kg_c0;            // largest allowed i
  if (lim < 0) return 0u;
  if (lim >= 31) return keep_word;
  return keep_word & ((2u << lim) - 1u);
}
// key-padding flag of key kc (thread tid < 128 owns key kg0 + tid of a block): loaded one block ahead so the global
// load latency is off the critical path
__device__ __forceinline__ bool load_keep(const AttnParams& p, int b, int kc, int tid) {
  if (tid >= AT) return false;
  bool kp = kc < p.Sk;
  if (kp && p.key_keep) kp = p.key_keep[(long long)b * p.Sk + kc] != 0;
  return kp;
}
__device__ __forceinline__ void publish_keep_bits(uint32_t* s_bits, bool kp, int tid) {
  if (tid < AT) {
    const uint32_t w = __ballot_sync(0xffffffffu, kp);
    if ((tid & 31) == 0) s_bits[tid >> 5] = w;
  }
}
// compute threads 0..127 publish the key-padding bitmap of key block kg0 as four 32-bit words
__device__ __forceinline__ void build_keep_bits(uint32_t* s_bits, const AttnParams& p, int b, int kg0, int tid) {
  if (tid < AT) {
    const int kc = kg0 + tid;
    bool kp = kc < p.Sk;
    if (kp && p.key_keep) kp = p.key_keep[(long long)b * p.Sk + kc] != 0;
    const uint32_t w = __ballot_sync(0xffffffffu, kp);
    if ((tid & 31) == 0) s_bits[tid >> 5] = w;
  }
}

// Tile coordinates of this CTA.  Causal problems have tiles of very different length (1..Sk/128 key blocks); CTAs are
// dispatched in linear block order, so the tile index along the sequence is made the SLOWEST coordinate there: all heads'
// longest tiles are issued first, the shortest last (longest-processing-time order; with the sequence index fastest the
// last wave mixed long and short tiles).  xt counts from the longest tile for the caller to map.
#ifndef PB_ATTN_LPT
#define PB_ATTN_LPT 1
#endif
__device__ __forceinline__ void tile_coords(bool causal, int& xt, int& h, int& b) {
  if (!PB_ATTN_LPT || !causal) { xt = blockIdx.x; h = blockIdx.y; b = blockIdx.z; return; }
  const int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const int hb = gridDim.y * gridDim.z;
  xt = lin / hb;
  const int r = lin - xt * hb;
  b = r / gridDim.y;
  h = r - b * gridDim.y;
}
struct Smem4 { uint8_t* t[6]; uint32_t a[6]; };
__device__ __forceinline__ void carve(uint8_t* raw, Smem4& s, int n) {
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  for (int i = 0; i < n; ++i) { s.t[i] = gen + i * TILE_BYTES; s.a[i] = base + i * TILE_BYTES; }
}

// ===================================================================================== forward
// The tensor pipe runs S = Q K^T two key blocks ahead of the softmax (three S buffers in TMEM, three-deep K ring), so
// that while the softmax warps exponentiate block j - a MUFU-bound stretch - the same instruction stream also takes
// the masked row maximum of block j+1 straight out of TMEM.  P(j) is written back (bf16, packed) over the S(j) columns
// it was computed from and feeds O += P V as the A operand from TMEM: no shared-memory tile, no proxy fence, no wait
// for the previous P V.  O accumulates in TMEM across key blocks and is only rescaled when the running row maximum
// grows by more than 2^8 ("lazy rescale").  Per block the softmax warps' critical path is
//   exchange max (one named barrier) -> [ld S(j), S(j+1) -> exp2 / max -> st P(j)] x 2 chunks -> arrive.
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 domain: probabilities stay <= 256 between rescales
constexpr int NSBUF = 3;                    // S buffers in TMEM / K ring depth
#ifdef PB_SINGLE_PHASE_BARRIERS             // round-1 behaviour, kept only for tools/attn_late_tile_demo.py
constexpr int NPB = 1;
#else
constexpr int NPB = NSBUF;                  // p_full / pv_done ring depth (see the comment at their declaration)
#endif
constexpr int FWD_MAX_SK = 8192;            // keys per sequence supported by the in-kernel key-padding bitmap

__global__ void __launch_bounds__(NTHREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap to, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 64) PB_TR(0, 63, 0);
  // pv_done is a ring of NSBUF barriers (P V(j) commits to pv_done[j % NSBUF]): the softmax warps only look at it when they
  // rescale O and at the end, i.e. they do NOT observe every phase, and a parity wait is only unambiguous when the phase
  // before the awaited one is known to be complete.  With one barrier, a late V tile (P V(j-2) still un-issued when the
  // warps ask for P V(j-1)) made the wait return immediately: O was read / rescaled early and the CTA could reach its
  // TMEM dealloc with MMAs in flight - the cold-start "unspecified launch failure" of round 1.  With the ring, the phase
  // before P V(j-1) on its barrier is P V(j-4), which has retired once S(j) is complete (in-order tensor pipe).
  // p_full is a ring for the mirror-image reason: S runs two blocks ahead, so the softmax warps can finish blocks j, j+1
  // and j+2 while the MMA thread still waits for the V tile of block j; with a single barrier its next parity wait then
  // aliased a later phase, and near the end of the key loop that phase never completes (deadlock -> bounded-spin trap).
  __shared__ __align__(8) uint64_t q_full, k_full[NSBUF], k_empty[NSBUF], v_full[2], v_empty[2], s_full[NSBUF], p_full[NPB], pv_done[NPB], mma_drain;
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t s_keep[FWD_MAX_SK / 32];   // key-padding bitmap of the whole key sequence (built once)
  __shared__ float s_red[2][2][AT];
  Smem4 sm;
  carve(smem_raw, sm, 6);  // 0 Q, 1-3 K ring, 4-5 V ring
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal: query blocks near the end of the sequence visit the most key blocks - schedule them first
  int xt, h, b;
  tile_coords(p.causal != 0, xt, h, b);
  const int qb = p.causal ? (int)gridDim.x - 1 - xt : xt;
  const int q0 = qb * AT;
  int nkb = (p.Sk + AT - 1) / AT;
  if (p.causal) nkb = min(nkb, qb + 1);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < NSBUF; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&s_full[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    mbar_init(&mma_drain, 1);
    for (int i = 0; i < NPB; ++i) { mbar_init(&pv_done[i], 1); mbar_init(&p_full[i], NCOMPUTE); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // global memory of the preceding kernel is visible from here on
  if (threadIdx.x == 64) PB_TR(0, 63, 1);
  const uint32_t tmem = tmem_base_smem;
  const uint32_t tS0 = tmem, tO = tmem + NSBUF * AT;   // S buffers at columns 0 / 128 / 256, O at 384

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&q_full, TILE_BYTES);
      load_tile(sm.t[0], &tq, &q_full, q0, h, b);
      // K runs two blocks ahead of V: S(j+2) is issued while block j is in the softmax, and a V slot is only released by
      // P V(j-2) - one producer thread must not let the K loads queue behind a V slot wait
      for (int i = 0; i < nkb + 2; ++i) {
        if (i < nkb) {
          const int ks = i % NSBUF;
          mbar_wait(&k_empty[ks], ((uint32_t)(i / NSBUF) & 1) ^ 1);
          chaos_delay(p.dbg_delay >> 2, 2 * i);
          mbar_expect_tx(&k_full[ks], TILE_BYTES);
          load_tile(sm.t[1 + ks], &tk, &k_full[ks], i * AT, h, b);
        }
        if (i >= 2) {
          const int j = i - 2, vs = j & 1;
          mbar_wait(&v_empty[vs], ((uint32_t)(j >> 1) & 1) ^ 1);
          chaos_delay(p.dbg_delay, 2 * j + 1);        // late V tiles are what the p_full / pv_done rings exist for
          mbar_expect_tx(&v_full[vs], TILE_BYTES);
          load_tile(sm.t[4 + vs], &tv, &v_full[vs], j * AT, h, b);
        }
      }
      // producer tail: every slot release (tcgen05.commit arrival on k_empty / v_empty) has landed before the CTA exits
      for (int i = nkb; i < nkb + NSBUF; ++i) mbar_wait(&k_empty[i % NSBUF], ((uint32_t)(i / NSBUF) & 1) ^ 1);
      for (int j = nkb; j < nkb + 2; ++j) mbar_wait(&v_empty[j & 1], ((uint32_t)(j >> 1) & 1) ^ 1);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      auto issue_s = [&](int i) {
        const int sb = i % NSBUF;
        mbar_wait(&k_full[sb], (uint32_t)(i / NSBUF) & 1);
        tc_fence_after();
        // buffer sb held S(i-3) / P(i-3): read by P V(i-3), which precedes this MMA in the in-order tensor pipe
        mma_tile<false, false>(tS0 + sb * AT, sm.a[0], sm.a[1 + sb], false);
        umma_commit(&s_full[sb]);
        umma_commit(&k_empty[sb]);
      };
      mbar_wait(&q_full, 0);
      for (int i = 0; i < 2 && i < nkb; ++i) issue_s(i);
      constexpr uint32_t idesc_pv = make_idesc_bf16(AT, AT, 0, 1);
      for (int j = 0; j < nkb; ++j) {
        if (j + 2 < nkb) { PB_TR(0, j, 0); issue_s(j + 2); PB_TR(0, j, 1); }
        mbar_wait(&p_full[j % NPB], (uint32_t)(j / NPB) & 1);
        mbar_wait(&v_full[j & 1], (uint32_t)(j >> 1) & 1);
        tc_fence_after();
        PB_TR(0, j, 2);
        // O += P V: A = P from TMEM (packed bf16 over S buffer j%3: key half 0 at columns 0-31, half 1 at 64-95),
        // B = V tile as MN-major operand
        const uint32_t tP = tS0 + (j % NSBUF) * AT, vt = sm.a[4 + (j & 1)];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ts(tO, tP + (kk >> 2) * 64 + (kk & 3) * 8, desc_mnmajor(vt, kk), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&pv_done[j % NPB]);
        umma_commit(&v_empty[j & 1]);
        PB_TR(0, j, 3);
      }
      // no MMA / commit of this CTA is in flight when TMEM is released and the CTA exits
      umma_commit(&mma_drain);
      mbar_wait(&mma_drain, 0);
    }
  } else {
    const int cw = warp - 2;                 // 0..7
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int hf = cw >> 2;                  // column half: S columns / O columns [64*hf, 64*hf + 64)
    const int r = quad * 32 + lane;          // tile row = TMEM lane
    const int tid = threadIdx.x - 64;        // 0..255
    const int qg = q0 + r;                   // global query index
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const bool causal = p.causal != 0;
    float m_used = -INFINITY, l = 0.f;       // scaling reference of the accumulator / partial row sum of this half
    const int trole = (lane == 0 && (warp == 2 || warp == 6)) ? (warp == 2 ? 1 : 2) : -1;
    // key-padding bitmap of all key blocks this CTA visits: one ballot per 32 keys, one barrier for the whole kernel
    for (int k0 = 0; k0 < nkb * AT; k0 += NCOMPUTE) {
      const int kc = k0 + tid;
      bool kp = kc < p.Sk;
      if (kp && p.key_keep) kp = p.key_keep[(long long)b * p.Sk + kc] != 0;
      const uint32_t w = __ballot_sync(0xffffffffu, kp);
      if (lane == 0) s_keep[kc >> 5] = w;
    }
    compute_bar_sync();
    // the two warps that share a row quadrant (column halves 0 / 1) exchange their row maxima through a 64-thread barrier
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory"); };
    uint32_t msk_cur[2], msk_next[2] = {0u, 0u};
#pragma unroll
    for (int c = 0; c < 2; ++c) msk_cur[c] = chunk_mask(s_keep[hf * 2 + c], causal, qg, hf * 64 + c * 32);
    // row maximum of block 0
    float m_blk;
    {
      mbar_wait(&s_full[0], 0);
      tc_fence_after();
      float bm = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tS0 + lane_addr + hf * 64 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) bm = fmaxf(bm, ((msk_cur[c] >> i) & 1u) ? __uint_as_float(v[i]) : -INFINITY);
      }
      s_red[0][hf][r] = bm * sl2;
      pair_sync();
      m_blk = fmaxf(s_red[0][0][r], s_red[0][1][r]);
    }
    for (int j = 0; j < nkb; ++j) {
      if (trole > 0) PB_TR(trole, j, 0);
      const bool has_next = j + 1 < nkb;
      const uint32_t tS = tS0 + (j % NSBUF) * AT, tSn = tS0 + ((j + 1) % NSBUF) * AT;
      // lazy rescale: keep the old reference unless the maximum grew by more than 2^8
      float f = 1.0f;
      bool need = false;
      if (m_used == -INFINITY) {
        m_used = m_blk;                                   // nothing non-zero accumulated so far
      } else if (m_blk > m_used + RESCALE_THRESHOLD) {
        f = ex2(m_used - m_blk);
        m_used = m_blk;
        need = true;
      }
      if (__any_sync(0xffffffffu, need)) {
        mbar_wait(&pv_done[(j - 1) % NPB], (uint32_t)((j - 1) / NPB) & 1);   // j > 0 here: the previous P V must have landed in O
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tO + lane_addr + hf * 64 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          tmem_st32(tO + lane_addr + hf * 64 + c * 32, v);
        }
        tmem_st_wait();
        l *= f;
      }
      const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
      if (has_next) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
          msk_next[c] = chunk_mask(s_keep[(j + 1) * 4 + hf * 2 + c], causal, qg, (j + 1) * AT + hf * 64 + c * 32);
        mbar_wait(&s_full[(j + 1) % NSBUF], (uint32_t)((j + 1) / NSBUF) & 1);
        tc_fence_after();
      }
      if (trole > 0) PB_TR(trole, j, 1);
      float rs0 = 0.f, rs1 = 0.f, bmn = -INFINITY;   // (row sums of the unrounded probabilities: two accumulators for ILP)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], nv[32];
        tmem_ld32(tS + lane_addr + hf * 64 + c * 32, sv);
        if (has_next) tmem_ld32(tSn + lane_addr + hf * 64 + c * 32, nv);
        tmem_ld_wait();
        uint32_t pk[16];
        const bool all_cur = msk_cur[c] == 0xffffffffu;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // masked entries are -inf -> ex2 gives 0
          float x0 = __uint_as_float(sv[2 * i]), x1 = __uint_as_float(sv[2 * i + 1]);
          if (!all_cur) {
            x0 = ((msk_cur[c] >> (2 * i)) & 1u) ? x0 : -INFINITY;
            x1 = ((msk_cur[c] >> (2 * i + 1)) & 1u) ? x1 : -INFINITY;
          }
          const float e0 = ex2(fmaf(x0, sl2, neg_m)), e1 = ex2(fmaf(x1, sl2, neg_m));
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
          rs0 += e0;
          rs1 += e1;
        }
        if (has_next) {                      // row maximum of the next block, hidden under the exponentials
          if (msk_next[c] == 0xffffffffu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) bmn = fmaxf(bmn, __uint_as_float(nv[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) bmn = fmaxf(bmn, ((msk_next[c] >> i) & 1u) ? __uint_as_float(nv[i]) : -INFINITY);
          }
        }
        // P chunk (32 keys = 16 packed columns) over S columns this thread has already consumed
        tmem_st16(tS + lane_addr + hf * 64 + c * 16, pk);
      }
      l += rs0 + rs1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[j % NPB]);
      if (trole > 0) PB_TR(trole, j, 2);
      if (has_next) {
        s_red[(j + 1) & 1][hf][r] = bmn * sl2;
        pair_sync();
        m_blk = fmaxf(s_red[(j + 1) & 1][0][r], s_red[(j + 1) & 1][1][r]);
        msk_cur[0] = msk_next[0]; msk_cur[1] = msk_next[1];
      }
      if (trole > 0) PB_TR(trole, j, 3);
    }
    mbar_wait(&pv_done[(nkb - 1) % NPB], (uint32_t)((nkb - 1) / NPB) & 1);
    tc_fence_after();
    if (threadIdx.x == 64) PB_TR(0, 63, 2);
    compute_bar_sync();
    s_red[0][hf][r] = l;
    compute_bar_sync();
    l = s_red[0][0][r] + s_red[0][1][r];
    const float inv = l > 0.f ? 1.f / l : 0.f;
    // O rows -> the Q tile (every S MMA has retired) -> two bulk tensor stores
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tO + lane_addr + hf * 64 + c * 32, v);
      tmem_ld_wait();
      stage_row_chunk(sm.t[0], r, hf, c, v, inv);
    }
    fence_proxy_async_smem();
    compute_bar_sync();
    if (tid == 0) { store_tile_tma(&to, sm.t[0], q0, h, b); bulk_wait_read<0>(); }
    if (qg < p.Sq && hf == 0) p.lse[((long long)b * p.H + h) * p.Sq + qg] = (l > 0.f) ? (m_used + log2f(l)) : INFINITY;
    if (threadIdx.x == 64) PB_TR(0, 63, 3);
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
  if (threadIdx.x == 64) PB_TR(0, 63, 4);
}

// ===================================================================================== forward, two Q tiles per CTA
// The production forward kernel.  One CTA owns TWO consecutive 128-query tiles (a, b) of a (batch, head) and one softmax
// warpgroup per tile (4 warps, thread = row, no exchange between threads); K / V tiles are shared by the two tiles and
// double-buffered.  The single MMA thread alternates between the tiles,
//     ... P V_a(j), S_a(j+1), P V_b(j), S_b(j+1), P V_a(j+1), ...
// so while warpgroup a exponentiates S_a(j+1) - a MUFU-bound stretch - the tensor pipe runs tile b's products and vice
// versa (TMEM: S_a | S_b | O_a | O_b = 512 columns; P_t is written back, packed, over the S_t columns it was derived from).
// Per tile every hand-off alternates strictly, S_t(j) -> softmax_t(j) -> P V_t(j) -> S_t(j+1) (in-order tensor pipe), so no
// parity wait can alias a later phase, and s_full_t(j) is committed after P V_t(j-1): the lazy O rescale needs no barrier of
// its own.  K(j+2) / V(j+2) are requested when S_b(j) / P V_b(j) retire, two block periods before their first use (a TMA
// refill takes 2000-3000 cycles under load, tools/gpu_attn_trace2.py).  Causal: tile a visits one key block fewer than b.
#ifndef PB_F3_POLY_MASK
#define PB_F3_POLY_MASK 0x2492   // pairs 1, 4, 7, 10, 13 of every 16: the measured balance point of the MUFU and FMA pipes
#endif
constexpr int F3_THREADS = 64 + 256 + 64;   // TMA, MMA, 8 softmax warps, 2 register-donor warps

__global__ void __launch_bounds__(F3_THREADS, 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                 const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap to, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 64) PB_TR(0, 63, 0);
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2], o_full[2], stagger, mma_drain;
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t s_keep[FWD_MAX_SK / 32];
  Smem4 sm;
  carve(smem_raw, sm, 6);  // 0-1 Q_a Q_b, 2-3 K ring, 4-5 V ring
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal: query tiles near the end of the sequence visit the most key blocks - schedule them first
  int xt, h, b;
  tile_coords(p.causal != 0, xt, h, b);
  const int qp = p.causal ? (int)gridDim.x - 1 - xt : xt;
  const int q0 = qp * 2 * AT;
  const int nkb_all = (p.Sk + AT - 1) / AT;
  int nk[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    nk[t] = p.causal ? min(nkb_all, 2 * qp + t + 1) : nkb_all;
    if (q0 + t * AT >= p.Sq) nk[t] = 0;                 // tile b beyond the sequence
  }
  const int na = nk[0], nb = max(nk[0], nk[1]);        // (na >= 1 always; nk[1] is 0 or >= na)

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1); mbar_init(&mma_drain, 1); mbar_init(&stagger, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  if (threadIdx.x == 64) PB_TR(0, 63, 1);
  const uint32_t tmem = tmem_base_smem;

  // Register budget (setmaxnreg, balanced per SM sub-partition = warp index mod 4): 12 warps are launched at 168 registers;
  // on every sub-partition one warp that needs few (TMA producer, MMA issuer, two idle donor warps) hands registers back and
  // the two softmax warps there grow to 224, so that a softmax thread keeps its whole S row (128 values) in registers next
  // to the packed P chunk it builds.
  if (warp == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (elect_one()) {
      mbar_expect_tx(&q_full, 2 * TILE_BYTES);
      load_tile(sm.t[0], &tq, &q_full, q0, h, b);
      load_tile(sm.t[1], &tq, &q_full, q0 + AT, h, b);
      // slot release order: K(j) [after S_b(j-2)], V(j) [after P V_b(j-2)], K(j+1) [after S_b(j-1)], ...
      for (int j = 0; j < nb; ++j) {
        const int s = j & 1;
        const uint32_t ph = ((uint32_t)(j >> 1) & 1) ^ 1;
        mbar_wait(&k_empty[s], ph);
        chaos_delay(p.dbg_delay >> 2, 2 * j);
        mbar_expect_tx(&k_full[s], TILE_BYTES);
        load_tile(sm.t[2 + s], &tk, &k_full[s], j * AT, h, b);
        mbar_wait(&v_empty[s], ph);
        chaos_delay(p.dbg_delay, 2 * j + 1);
        mbar_expect_tx(&v_full[s], TILE_BYTES);
        load_tile(sm.t[4 + s], &tv, &v_full[s], j * AT, h, b);
      }
      // producer tail: every slot release (tcgen05.commit arrival) has landed before the CTA exits
      for (int j = nb; j < nb + 2; ++j) {
        mbar_wait(&k_empty[j & 1], ((uint32_t)(j >> 1) & 1) ^ 1);
        mbar_wait(&v_empty[j & 1], ((uint32_t)(j >> 1) & 1) ^ 1);
      }
    }
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (elect_one()) {
      constexpr uint32_t idesc_pv = make_idesc_bf16(AT, AT, 0, 1);
      const bool two = nk[1] > 0;
      auto issue_s = [&](int t, int j) {           // S_t(j) = Q_t K(j)^T; overwrites P_t(j-1), read by the P V_t(j-1) issued before it
        mma_tile<false, false>(tmem + t * AT, sm.a[t], sm.a[2 + (j & 1)], false);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {          // O_t += P_t(j) V(j): A = packed bf16 P over the S_t columns, B = V tile MN-major
        const uint32_t tP = tmem + t * AT, tO = tmem + 2 * AT + t * AT, vt = sm.a[4 + (j & 1)];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ts(tO, tP + kk * 8, desc_mnmajor(vt, kk), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
      };
      mbar_wait_spin(&q_full, 0);
      mbar_wait_spin(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      if (two) {
        // Tile b starts about half a block period after tile a (warpgroup a signals `stagger` part-way through its first
        // block): the two softmax warpgroups then run out of phase - one exponentiates while the other waits for its
        // products - instead of contending for the MUFU pipe and then leaving the tensor pipe idle together.  The phase
        // offset persists: both tiles have the same period.
        mbar_wait_spin(&stagger, 0);
        issue_s(1, 0);
      }
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nb; ++j) {
        const int s = j & 1, sn = (j + 1) & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1, phn = (uint32_t)((j + 1) >> 1) & 1;
        bool v_ok = false, k_ok = false;
        if (j < na) {
          PB_TR(0, j, 0);
          mbar_wait_spin(&p_full[0], (uint32_t)j & 1);
          mbar_wait_spin(&v_full[s], ph); v_ok = true;
          tc_fence_after();
          PB_TR(0, j, 1);
          issue_pv(0, j);
          if (j + 1 < na) {
            mbar_wait_spin(&k_full[sn], phn); k_ok = true;
            tc_fence_after();
            issue_s(0, j + 1);
          } else {
            umma_commit(&o_full[0]);
          }
          if (!two) {                                // single-tile CTA (sequence tail): tile a releases the slots
            umma_commit(&v_empty[s]);
            if (j + 1 < na) umma_commit(&k_empty[sn]);
          }
          PB_TR(0, j, 2);
        }
        if (two) {
          mbar_wait_spin(&p_full[1], (uint32_t)j & 1);
          if (!v_ok) mbar_wait_spin(&v_full[s], ph);
          tc_fence_after();
          PB_TR(0, j, 3);
          issue_pv(1, j);
          umma_commit(&v_empty[s]);
          if (j + 1 < nb) {
            if (!k_ok) mbar_wait_spin(&k_full[sn], phn);
            tc_fence_after();
            issue_s(1, j + 1);
            umma_commit(&k_empty[sn]);
          } else {
            umma_commit(&o_full[1]);
          }
          PB_TR(0, j, 4);
        }
      }
      // no MMA / commit of this CTA is in flight when TMEM is released and the CTA exits
      umma_commit(&mma_drain);
      mbar_wait(&mma_drain, 0);
    }
  } else if (warp >= 10) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");     // donor warps: nothing else to do
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 2) >> 2;           // tile / warpgroup
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;          // tile row = TMEM lane
    const int tid = threadIdx.x - 64;        // 0..255
    const int qg = q0 + t * AT + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t tS = tmem + t * AT + lane_addr, tO = tmem + 2 * AT + t * AT + lane_addr;
    const float sl2 = p.scale * LOG2E;
    const bool causal = p.causal != 0;
    const int nkt = nk[t];
    float m_used = -INFINITY, l = 0.f;
    const int trole = (lane == 0 && (warp == 2 || warp == 6)) ? (warp == 2 ? 1 : 2) : -1;
    for (int k0 = 0; k0 < nb * AT; k0 += 256) {
      const int kc = k0 + tid;
      bool kp = kc < p.Sk;
      if (kp && p.key_keep) kp = p.key_keep[(long long)b * p.Sk + kc] != 0;
      const uint32_t w = __ballot_sync(0xffffffffu, kp);
      if (lane == 0 && kc < FWD_MAX_SK) s_keep[kc >> 5] = w;
    }
    compute_bar_sync();
    for (int j = 0; j < nkt; ++j) {
      uint32_t msk[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) msk[c] = chunk_mask(s_keep[j * 4 + c], causal, qg, j * AT + c * 32);
      if (trole > 0) PB_TR(trole, j, 0);
      mbar_wait(&s_full[t], (uint32_t)j & 1);
      tc_fence_after();
      if (trole > 0) PB_TR(trole, j, 1);
      // the whole S row (128 fp32 values) is read from TMEM once and stays in registers for both passes
      uint32_t v0[32], v1[32], v2[32], v3[32];
      tmem_ld32(tS, v0); tmem_ld32(tS + 32, v1); tmem_ld32(tS + 64, v2); tmem_ld32(tS + 96, v3);
      tmem_ld_wait();
      // pass 1: masked row maximum
      float bm0 = -INFINITY, bm1 = -INFINITY;
      auto rmax = [&](const uint32_t (&v)[32], uint32_t m) {
        if (m == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) { bm0 = fmaxf(bm0, __uint_as_float(v[i])); bm1 = fmaxf(bm1, __uint_as_float(v[i + 1])); }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            bm0 = fmaxf(bm0, ((m >> i) & 1u) ? __uint_as_float(v[i]) : -INFINITY);
            bm1 = fmaxf(bm1, ((m >> (i + 1)) & 1u) ? __uint_as_float(v[i + 1]) : -INFINITY);
          }
        }
      };
      rmax(v0, msk[0]); rmax(v1, msk[1]); rmax(v2, msk[2]); rmax(v3, msk[3]);
      const float m_blk = fmaxf(bm0, bm1) * sl2;
      if (trole > 0) PB_TR(trole, j, 2);
      // lazy rescale: keep the old reference unless the maximum grew by more than 2^8
      float f = 1.0f;
      bool need = false;
      if (m_used == -INFINITY) {
        m_used = m_blk;
      } else if (m_blk > m_used + RESCALE_THRESHOLD) {
        f = ex2(m_used - m_blk);
        m_used = m_blk;
        need = true;
      }
      if (__any_sync(0xffffffffu, need)) {     // (P V_t(j-1) has retired: s_full_t(j) was committed after it)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(tO + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          tmem_st32(tO + c * 32, v);
        }
        l *= f;
      }
      const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
      if (trole > 0) PB_TR(trole, j, 3);
      // pass 2: exponentials; P chunk c (32 keys = 16 packed columns) is written back over the S columns
      // (an unmasked and a masked instance of the chunk body behind a warp-uniform branch: masked key blocks - padding, the
      // causal diagonal - are the minority, and predicated-off mask instructions would still cost their issue slots)
      float2 rs = make_float2(0.f, 0.f);
      const float2 sl2v = make_float2(sl2, sl2), negv = make_float2(neg_m, neg_m);
      auto expo = [&](auto masked, const uint32_t (&sv)[32], uint32_t m, int c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 x = make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1]));
          if constexpr (decltype(masked)::value) {
            x.x = ((m >> (2 * i)) & 1u) ? x.x : -INFINITY;      // masked entries are -inf -> ex2 gives 0
            x.y = ((m >> (2 * i + 1)) & 1u) ? x.y : -INFINITY;
          }
          const float2 y = ffma2(x, sl2v, negv);
          float2 e;
          if constexpr (!decltype(masked)::value) {
            if ((PB_F3_POLY_MASK >> i) & 1) e = exp2_poly2(y);   // this share of the pairs on the FMA pipe
            else e = make_float2(ex2(y.x), ex2(y.y));
          } else {
            e = make_float2(ex2(y.x), ex2(y.y));                 // (-inf must give exactly 0)
          }
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(e.x, e.y);
          pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
          rs = fadd2(rs, e);
        }
        tmem_st16(tS + c * 16, pk);
      };
      const bool unmasked = __all_sync(0xffffffffu, (msk[0] & msk[1] & msk[2] & msk[3]) == 0xffffffffu);
      if (unmasked) {
        expo(std::false_type{}, v0, msk[0], 0);
        if (t == 0 && j == 0) mbar_arrive(&stagger);
        expo(std::false_type{}, v1, msk[1], 1); expo(std::false_type{}, v2, msk[2], 2); expo(std::false_type{}, v3, msk[3], 3);
      } else {
        expo(std::true_type{}, v0, msk[0], 0);
        if (t == 0 && j == 0) mbar_arrive(&stagger);
        expo(std::true_type{}, v1, msk[1], 1); expo(std::true_type{}, v2, msk[2], 2); expo(std::true_type{}, v3, msk[3], 3);
      }
      const float rs0 = rs.x, rs1 = rs.y;
      l += rs0 + rs1;
      tmem_st_wait();
      if (trole > 0) PB_TR(trole, j, 4);
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      if (trole > 0) PB_TR(trole, j, 5);
    }
    if (nkt > 0) {
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      if (threadIdx.x == 64) PB_TR(0, 63, 2);
      const float inv = l > 0.f ? 1.f / l : 0.f;
      // O rows -> this tile's Q buffer (every S_t MMA has retired) -> two bulk tensor stores
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tO + c * 32, v);
        tmem_ld_wait();
        stage_row_chunk(sm.t[t], r, c >> 1, c & 1, v, inv);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(2 + t) : "memory");
      if ((tid & 127) == 0) { store_tile_tma(&to, sm.t[t], q0 + t * AT, h, b); bulk_wait_read<0>(); }
      if (qg < p.Sq) p.lse[((long long)b * p.H + h) * p.Sq + qg] = (l > 0.f) ? (m_used + log2f(l)) : INFINITY;
      if (threadIdx.x == 64) PB_TR(0, 63, 3);
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ===================================================================================== backward: dK, dV
// Transposed formulation so that 64-query blocks keep every MMA at M = 128:  S^T = K Q^T and dP^T = V dO^T are
// [128 keys x 64 queries] (TMEM lane = key); P^T / dS^T are written back (bf16, packed) over the S^T / dP^T columns they
// were derived from and feed dV += P^T dO, dK += dS^T Q as A operands from TMEM, with the 64-row Q / dO tiles as MN-major
// B operands - no shared-memory round trip.  Q/dO stream through a 5-deep ring (a TMA refill takes > 1500 cycles under
// load), S^T/dP^T are double-buffered: TMA, the tensor pipe and the softmax warps overlap instead of taking turns (the
// v1 kernel spent 43% of its samples waiting on the S/dP barrier, profiles/r1_summary.md).
constexpr int QB = 64;                          // queries per block in this kernel
constexpr int QT_BYTES = QB * AT * 2;           // 16 KB: [64 queries x 128 head dims], two 64-column halves of 8 KB
constexpr int NQ = 5;                           // Q/dO ring depth

__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                    const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                    const __grid_constant__ CUtensorMap tdk, const __grid_constant__ CUtensorMap tdv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t kv_full, qdo_full[NQ], qdo_empty[NQ], sdp_full[2], pds_full[2], acc_full, mma_drain;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_L[2][QB], s_D[2][QB];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // layout: K 32K | V 32K | Q ring NQ x 16K | dO ring NQ x 16K
  const uint32_t aK = base, aV = base + TILE_BYTES, aQ = base + 2 * TILE_BYTES, adO = aQ + NQ * QT_BYTES;
  uint8_t* gK = gen; uint8_t* gV = gen + TILE_BYTES; uint8_t* gQ = gen + 2 * TILE_BYTES; uint8_t* gdO = gQ + NQ * QT_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int kb, h, b;
  tile_coords(p.causal != 0, kb, h, b);     // key block 0 sees every query block: longest first
  const int k0 = kb * AT;
  const int nqb = (p.Sq + QB - 1) / QB;
  const int qb0 = p.causal ? (k0 / QB) : 0;   // causal: first 64-query block that can see key k0
  const int niter = nqb - qb0;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo); }
  if (warp == 1 && lane == 0) {
    mbar_init(&kv_full, 1); mbar_init(&acc_full, 1); mbar_init(&mma_drain, 1);
    for (int i = 0; i < NQ; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&pds_full[i], NCOMPUTE); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // global memory of the preceding kernel is visible from here on
  const uint32_t tmem = tmem_base_smem;
  // TMEM columns: S^T[2] at 0/64, dP^T[2] at 128/192, dV at 256, dK at 384
  const uint32_t tS0 = tmem, tdP0 = tmem + 128, tdV = tmem + 256, tdK = tmem + 384;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&kv_full, 2 * TILE_BYTES);
      load_tile(gK, &tk, &kv_full, k0, h, b);
      load_tile(gV, &tv, &kv_full, k0, h, b);
      for (int it = 0; it < niter; ++it) {
        const int st = it % NQ;
        const int q0 = (qb0 + it) * QB;
        mbar_wait(&qdo_empty[st], ((uint32_t)(it / NQ) & 1) ^ 1);
        chaos_delay(p.dbg_delay, it);
        mbar_expect_tx(&qdo_full[st], 2 * QT_BYTES);
        tma_load_4d(gQ + st * QT_BYTES, &tq, &qdo_full[st], 0, q0, h, b);
        tma_load_4d(gQ + st * QT_BYTES + QT_BYTES / 2, &tq, &qdo_full[st], 64, q0, h, b);
        tma_load_4d(gdO + st * QT_BYTES, &tdo, &qdo_full[st], 0, q0, h, b);
        tma_load_4d(gdO + st * QT_BYTES + QT_BYTES / 2, &tdo, &qdo_full[st], 64, q0, h, b);
      }
      // producer tail: every ring-slot release has landed before the CTA exits
      for (int it = niter; it < niter + NQ; ++it) mbar_wait(&qdo_empty[it % NQ], ((uint32_t)(it / NQ) & 1) ^ 1);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(AT, QB, 0, 0);     // [128 keys x 64 q], both operands K-major
      constexpr uint32_t idesc_a = make_idesc_bf16(AT, AT, 0, 1);     // [128 keys x 128 hd], B = Q / dO tile MN-major
      auto issue_sdp = [&](int it) {
        const int st = it & 1;
        const uint32_t qt = aQ + (it % NQ) * QT_BYTES, ot = adO + (it % NQ) * QT_BYTES;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {   // contraction over head_dim
          const uint64_t kd = make_smem_desc_sw128(aK + (kk >> 2) * HALF_BYTES + (kk & 3) * 32, 16, 1024);
          const uint64_t qd = make_smem_desc_sw128(qt + (kk >> 2) * (QT_BYTES / 2) + (kk & 3) * 32, 16, 1024);
          umma_bf16(tS0 + st * QB, kd, qd, idesc_s, kk > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t vd = make_smem_desc_sw128(aV + (kk >> 2) * HALF_BYTES + (kk & 3) * 32, 16, 1024);
          const uint64_t od = make_smem_desc_sw128(ot + (kk >> 2) * (QT_BYTES / 2) + (kk & 3) * 32, 16, 1024);
          umma_bf16(tdP0 + st * QB, vd, od, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&sdp_full[st]);
      };
      mbar_wait(&kv_full, 0);
      if (niter > 0) {
        mbar_wait(&qdo_full[0], 0);
        tc_fence_after();
        issue_sdp(0);
      }
      for (int it = 0; it < niter; ++it) {
        const int st = it & 1;
        if (it + 1 < niter) {
          mbar_wait(&qdo_full[(it + 1) % NQ], (uint32_t)((it + 1) / NQ) & 1);
          tc_fence_after();
          issue_sdp(it + 1);             // its TMEM buffers held P^T / dS^T of block it-1: read by the dV / dK MMAs
        }                                // of block it-1, which precede these in the in-order tensor pipe
        mbar_wait(&pds_full[st], (uint32_t)(it >> 1) & 1);
        tc_fence_after();
        const uint32_t qt = aQ + (it % NQ) * QT_BYTES, ot = adO + (it % NQ) * QT_BYTES;
        // A operands from TMEM: query 16*kk.. of the block = packed columns (kk >> 1) * 32 + (kk & 1) * 8 of buffer st
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {   // contraction over the 64 queries of the block
          const uint32_t acol = (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8);
          const uint64_t bo = make_smem_desc_sw128(ot + kk * 2048, QT_BYTES / 2, 1024);     // dO: MN-major
          umma_bf16_ts(tdV, tS0 + st * QB + acol, bo, idesc_a, (it > 0 || kk > 0) ? 1u : 0u);
          const uint64_t bq = make_smem_desc_sw128(qt + kk * 2048, QT_BYTES / 2, 1024);     // Q: MN-major
          umma_bf16_ts(tdK, tdP0 + st * QB + acol, bq, idesc_a, (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&qdo_empty[it % NQ]);    // Q / dO ring slot reusable
      }
      umma_commit(&acc_full);
      umma_commit(&mma_drain);
      mbar_wait(&mma_drain, 0);
    }
  } else {
    const int cw = warp - 2;
    const int quad = warp & 3;
    const int hf = cw >> 2;                  // 32-query half of the 64-query block
    const int r = quad * 32 + lane;          // key row = TMEM lane
    const int tid = threadIdx.x - 64;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const int kg = k0 + r;
    bool kkeep = kg < p.Sk;
    if (kkeep && p.key_keep) kkeep = p.key_keep[(long long)b * p.Sk + kg] != 0;
    const long long rbase = ((long long)b * p.H + h) * p.Sq;
    // per-query statistics (LSE, D) of block it+1 are staged in shared memory while block it is processed
    // staged as -LSE and -D * scale so that P = exp2(fma(S, scale*log2e, -LSE)) and dS = P * fma(dP, scale, -D*scale)
    auto load_L = [&](int q) { return (tid < QB && q < p.Sq) ? -p.lse[rbase + q] : -INFINITY; };
    auto load_D = [&](int q) { return (tid < QB && q < p.Sq) ? -p.dvec[rbase + q] * p.scale : 0.f; };
    float L_next = 0.f, D_next = 0.f;
    if (tid < QB) {
      s_L[0][tid] = load_L(qb0 * QB + tid); s_D[0][tid] = load_D(qb0 * QB + tid);
      L_next = load_L((qb0 + 1) * QB + tid); D_next = load_D((qb0 + 1) * QB + tid);
    }
    for (int it = 0; it < niter; ++it) {
      const int st = it & 1;
      const int qbase = (qb0 + it) * QB + hf * 32;     // query of column 0 of this thread's chunk
      compute_bar_sync();                    // publishes stats(it); orders reuse of the other stats buffer
      if (tid < QB) {
        s_L[st ^ 1][tid] = L_next; s_D[st ^ 1][tid] = D_next;
        L_next = load_L((qb0 + it + 2) * QB + tid); D_next = load_D((qb0 + it + 2) * QB + tid);
      }
      // bit i: key kg may be attended by query qbase + i  (key padding; causal: kg <= q)
      uint32_t msk = kkeep ? 0xffffffffu : 0u;
      if (p.causal) {
        const int lim = kg - qbase;          // columns i < lim are masked
        if (lim >= 32) msk = 0u;
        else if (lim > 0) msk &= ~((1u << lim) - 1u);
      }
      mbar_wait(&sdp_full[st], (uint32_t)(it >> 1) & 1);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld32(tS0 + st * QB + lane_addr + hf * 32, sv);
      tmem_ld32(tdP0 + st * QB + lane_addr + hf * 32, dv);
      tmem_ld_wait();
      // P^T / dS^T of this thread's 32 queries: 16 packed columns over the S^T / dP^T columns it has just consumed
      uint32_t pp[16], pd[16];
      const float4* L4 = reinterpret_cast<const float4*>(&s_L[st][hf * 32]);
      const float4* D4 = reinterpret_cast<const float4*>(&s_D[st][hf * 32]);
      if (msk == 0xffffffffu) {
        // unmasked chunk (the common case): packed fp32 arithmetic, a share of the exponentials on the FMA pipe
        const float2 sl2v = make_float2(sl2, sl2), scv = make_float2(p.scale, p.scale);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 l4 = L4[g], d4 = D4[g];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int i = g * 2 + t;         // pair index: queries 2i, 2i+1 of the chunk
            const float2 nl = t == 0 ? make_float2(l4.x, l4.y) : make_float2(l4.z, l4.w);
            const float2 nd = t == 0 ? make_float2(d4.x, d4.y) : make_float2(d4.z, d4.w);
            float2 e, dsv;
            if ((PB_BWD_POLY_MASK >> i) & 1) bwd_pair<true>(sv[2 * i], sv[2 * i + 1], dv[2 * i], dv[2 * i + 1], sl2v, nl, scv, nd, e, dsv);
            else bwd_pair<false>(sv[2 * i], sv[2 * i + 1], dv[2 * i], dv[2 * i + 1], sl2v, nl, scv, nd, e, dsv);
            const __nv_bfloat162 a2 = __floats2bfloat162_rn(e.x, e.y);
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(dsv.x, dsv.y);
            pp[i] = *reinterpret_cast<const uint32_t*>(&a2);
            pd[i] = *reinterpret_cast<const uint32_t*>(&b2);
          }
        }
      } else {
        float pr[32], ds[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 l4 = L4[g], d4 = D4[g];
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int i = g * 4 + t;
            const float pv = ((msk >> i) & 1u) ? ex2(fmaf(__uint_as_float(sv[i]), sl2, lv[t])) : 0.f;
            pr[i] = pv;
            ds[i] = pv * fmaf(__uint_as_float(dv[i]), p.scale, dd[t]);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __nv_bfloat162 a2 = __floats2bfloat162_rn(pr[2 * i], pr[2 * i + 1]);
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(ds[2 * i], ds[2 * i + 1]);
          pp[i] = *reinterpret_cast<const uint32_t*>(&a2);
          pd[i] = *reinterpret_cast<const uint32_t*>(&b2);
        }
      }
      tmem_st16(tS0 + st * QB + lane_addr + hf * 32, pp);
      tmem_st16(tdP0 + st * QB + lane_addr + hf * 32, pd);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pds_full[st]);
    }
    // epilogue: thread = (key row, 64-column half of head_dim)
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    // dV rows -> the V tile, dK rows -> the K tile (all S^T / dP^T MMAs have retired) -> bulk tensor stores
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t tacc = which == 0 ? tdV : tdK;
      uint8_t* tile = which == 0 ? gV : gK;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        if (niter > 0) {
          tmem_ld32(tacc + lane_addr + hf * 64 + c * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        stage_row_chunk(tile, r, hf, c, v, 1.0f);
      }
    }
    fence_proxy_async_smem();
    compute_bar_sync();
    if (tid == 0) { store_tile_tma(&tdv, gV, k0, h, b); store_tile_tma(&tdk, gK, k0, h, b); bulk_wait_read<0>(); }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ===================================================================================== backward: dQ
// Every tcgen05.mma (M = 128, K = 16) occupies the tensor pipe for max(~72, N/2) cycles (tools/micro/mma_bench.cu), so
// all products use N = 128: 128-key blocks, 24 MMAs per block.  S is double-buffered in TMEM, dP single-buffered but
// released as soon as the softmax warps hold it in registers, and dS never touches shared memory: it is written (bf16,
// packed) over the S buffer it was derived from and feeds dQ += dS K as the A operand from TMEM:
//   tensor pipe :  S(j+1)  dP(j+1)  dQ(j)   S(j+2) ...          softmax warps :  block j -> dS(j)   block j+1 ...
// K streams through a three-deep ring (a slot is held from S(j) to dQ(j), and a TMA refill takes ~1500 cycles), V through a
// two-deep ring (released after dP).
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                   const __grid_constant__ CUtensorMap tdq, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t qdo_full, k_full[3], k_empty[3], v_full[2], v_empty[2], s_full[2], dp_full, dp_free, ds_full,
      acc_full, mma_drain;
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t s_bits2[2][4];
  Smem4 sm;
  {
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    for (int i = 0; i < 6; ++i) { sm.t[i] = gen + i * TILE_BYTES; sm.a[i] = base + i * TILE_BYTES; }
  }
  // tiles: 0 Q, 1 dO, 2-4 K ring, 5-6 V ring
  constexpr int NK = 3;
  uint8_t* gV = sm.t[5];
  const uint32_t aV = sm.a[5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // causal: query blocks near the end of the sequence visit the most key blocks - schedule them first
  int xt, h, b;
  tile_coords(p.causal != 0, xt, h, b);
  const int qb = p.causal ? (int)gridDim.x - 1 - xt : xt;
  const int q0 = qb * AT;
  int nkb = (p.Sk + AT - 1) / AT;
  if (p.causal) nkb = min(nkb, qb + 1);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo); }
  if (warp == 1 && lane == 0) {
    mbar_init(&qdo_full, 1); mbar_init(&acc_full, 1); mbar_init(&dp_full, 1); mbar_init(&mma_drain, 1);
    mbar_init(&dp_free, NCOMPUTE); mbar_init(&ds_full, NCOMPUTE);
    for (int i = 0; i < NK; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); mbar_init(&s_full[i], 1); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // global memory of the preceding kernel is visible from here on
  const uint32_t tmem = tmem_base_smem;
  const uint32_t tS0 = tmem, tdP = tmem + 256, tdQ = tmem + 384;   // S buffers at columns 0 / 128

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&qdo_full, 2 * TILE_BYTES);
      load_tile(sm.t[0], &tq, &qdo_full, q0, h, b);
      load_tile(sm.t[1], &tdo, &qdo_full, q0, h, b);
      for (int j = 0; j < nkb; ++j) {
        const int ks = j % NK, vs = j & 1;
        mbar_wait(&k_empty[ks], ((uint32_t)(j / NK) & 1) ^ 1);
        chaos_delay(p.dbg_delay, 2 * j);
        mbar_expect_tx(&k_full[ks], TILE_BYTES);
        load_tile(sm.t[2 + ks], &tk, &k_full[ks], j * AT, h, b);
        mbar_wait(&v_empty[vs], ((uint32_t)(j >> 1) & 1) ^ 1);
        chaos_delay(p.dbg_delay, 2 * j + 1);
        mbar_expect_tx(&v_full[vs], TILE_BYTES);
        load_tile(gV + vs * TILE_BYTES, &tv, &v_full[vs], j * AT, h, b);
      }
      // producer tail: every ring-slot release has landed before the CTA exits
      for (int j = nkb; j < nkb + NK; ++j) mbar_wait(&k_empty[j % NK], ((uint32_t)(j / NK) & 1) ^ 1);
      for (int j = nkb; j < nkb + 2; ++j) mbar_wait(&v_empty[j & 1], ((uint32_t)(j >> 1) & 1) ^ 1);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(&qdo_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      mma_tile<false, false>(tS0, sm.a[0], sm.a[2], false);                 // S(0) = Q K(0)^T
      umma_commit(&s_full[0]);
      mbar_wait(&v_full[0], 0);
      tc_fence_after();
      mma_tile<false, false>(tdP, sm.a[1], aV, false);                      // dP(0) = dO V(0)^T
      umma_commit(&dp_full);
      umma_commit(&v_empty[0]);
      for (int j = 0; j < nkb; ++j) {
        const int st = j & 1, nx = st ^ 1;
        const uint32_t ph = (uint32_t)j & 1;
        if (j + 1 < nkb) {
          // dP first: its inputs (V ring, dP buffer released at the very start of softmax(j)) are ready long before K(j+1),
          // whose ring slot was only released by dQ(j-2) - issuing S(j+1) second gives that TMA refill half a block more
          mbar_wait(&dp_free, ph);                                            // softmax(j) holds dP(j) in registers
          mbar_wait(&v_full[nx], (uint32_t)((j + 1) >> 1) & 1);
          tc_fence_after();
          PB_TR(0, j, 0);
          mma_tile<false, false>(tdP, sm.a[1], aV + nx * TILE_BYTES, false);  // dP(j+1)
          umma_commit(&dp_full);
          umma_commit(&v_empty[nx]);
          mbar_wait(&k_full[(j + 1) % NK], (uint32_t)((j + 1) / NK) & 1);
          tc_fence_after();
          PB_TR(0, j, 1);
          // S buffer nx held S(j-1) / dS(j-1): read by dQ(j-1), which precedes this MMA in the in-order tensor pipe
          mma_tile<false, false>(tS0 + nx * AT, sm.a[0], sm.a[2 + (j + 1) % NK], false);   // S(j+1)
          umma_commit(&s_full[nx]);
#ifdef PB_TRACE_EXEC
          PB_TR(0, j, 4);
          mbar_wait_spin(&s_full[nx], (uint32_t)((j + 1) >> 1) & 1);
          PB_TR(0, j, 5);
#endif
        }
        mbar_wait(&ds_full, ph);
        tc_fence_after();
        PB_TR(0, j, 2);
        {
          // dQ += dS K: A = dS from TMEM (packed bf16 over S buffer st: key half 0 at columns 0-31, half 1 at 64-95),
          // B = K tile as MN-major operand
          constexpr uint32_t idesc = make_idesc_bf16(AT, AT, 0, 1);
          const uint32_t kt = sm.a[2 + j % NK];
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16_ts(tdQ, tS0 + st * AT + (kk >> 2) * 64 + (kk & 3) * 8, desc_mnmajor(kt, kk), idesc, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&k_empty[j % NK]);                                        // K ring slot reusable
        PB_TR(0, j, 3);
      }
      umma_commit(&acc_full);
      umma_commit(&mma_drain);
      mbar_wait(&mma_drain, 0);
    }
  } else {
    const int cw = warp - 2;
    const int quad = warp & 3;
    const int hf = cw >> 2;                  // 64-column half of the key block handled by this thread
    const int r = quad * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const int qg = q0 + r;
    const bool qok = qg < p.Sq;
    const long long ridx = ((long long)b * p.H + h) * p.Sq + qg;
    const float negL = qok ? -p.lse[ridx] : -INFINITY;
    const float nDs = qok ? -p.dvec[ridx] * p.scale : 0.f;
    publish_keep_bits(s_bits2[0], load_keep(p, b, tid, tid), tid);
    bool kp_next = load_keep(p, b, AT + tid, tid);
    for (int j = 0; j < nkb; ++j) {
      const int kg0 = j * AT, st = j & 1;
      const uint32_t ph = (uint32_t)j & 1;
      compute_bar_sync();                    // publishes bits(j); orders reuse of the other bitmap buffer
      const int trole = (lane == 0 && (warp == 2 || warp == 6)) ? (warp == 2 ? 1 : 2) : -1;
      if (trole > 0) PB_TR(trole, j, 0);
      uint32_t msk[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) msk[c] = qok ? chunk_mask(s_bits2[st][hf * 2 + c], p.causal != 0, qg, kg0 + hf * 64 + c * 32) : 0u;
      publish_keep_bits(s_bits2[st ^ 1], kp_next, tid);
      kp_next = load_keep(p, b, (j + 2) * AT + tid, tid);
      mbar_wait(&s_full[st], (uint32_t)(j >> 1) & 1);
      mbar_wait(&dp_full, ph);
      tc_fence_after();
      if (trole > 0) PB_TR(trole, j, 1);
      uint32_t dv[2][32];
      tmem_ld32(tdP + lane_addr + hf * 64, dv[0]);
      tmem_ld32(tdP + lane_addr + hf * 64 + 32, dv[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&dp_free);                 // dP(j) is in registers: the tensor pipe may overwrite it with dP(j+1)
      if (trole > 0) PB_TR(trole, j, 2);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32];
        tmem_ld32(tS0 + st * AT + lane_addr + hf * 64 + c * 32, sv);
        tmem_ld_wait();
        uint32_t pk[16];
        if (msk[c] == 0xffffffffu) {
          // unmasked chunk (the common case): packed fp32 arithmetic, a share of the exponentials on the FMA pipe
          const float2 sl2v = make_float2(sl2, sl2), scv = make_float2(p.scale, p.scale);
          const float2 nl = make_float2(negL, negL), nd = make_float2(nDs, nDs);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 e, dsv;
            if ((PB_BWD_POLY_MASK >> i) & 1) bwd_pair<true>(sv[2 * i], sv[2 * i + 1], dv[c][2 * i], dv[c][2 * i + 1], sl2v, nl, scv, nd, e, dsv);
            else bwd_pair<false>(sv[2 * i], sv[2 * i + 1], dv[c][2 * i], dv[c][2 * i + 1], sl2v, nl, scv, nd, e, dsv);
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(dsv.x, dsv.y);
            pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ((msk[c] >> (2 * i)) & 1u) ? ex2(fmaf(__uint_as_float(sv[2 * i]), sl2, negL)) : 0.f;
            const float p1 = ((msk[c] >> (2 * i + 1)) & 1u) ? ex2(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, negL)) : 0.f;
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(p0 * fmaf(__uint_as_float(dv[c][2 * i]), p.scale, nDs),
                                                            p1 * fmaf(__uint_as_float(dv[c][2 * i + 1]), p.scale, nDs));
            pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
          }
        }
        if (c == 0 && trole > 0) PB_TR(trole, j, 3);
        // dS chunk (32 keys = 16 packed columns) over the S columns this thread has already consumed
        tmem_st16(tS0 + st * AT + lane_addr + hf * 64 + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full);
      if (trole > 0) PB_TR(trole, j, 5);
    }
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    // dQ rows -> the Q tile (every S MMA has retired) -> two bulk tensor stores
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tdQ + lane_addr + hf * 64 + c * 32, v);
      tmem_ld_wait();
      stage_row_chunk(sm.t[0], r, hf, c, v, 1.0f);
    }
    fence_proxy_async_smem();
    compute_bar_sync();
    if (tid == 0) { store_tile_tma(&tdq, sm.t[0], q0, h, b); bulk_wait_read<0>(); }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// D[b,h,q] = sum_c dO[b,q,h,c] * O[b,q,h,c]      (16 lanes x 16 bytes per (b,q,h) row of 128, two rows in flight per thread)
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                            float* __restrict__ dvec, int B, int H, int Sq, long long ldo,
                                                            long long o_sb, long long lddo, long long do_sb) {
  pdl_entry();
  const int sub = threadIdx.x & 15;                                      // 16-byte piece of the row
  const long long total = (long long)B * Sq * H;
  const long long g0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;   // row group of this half-warp
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 4;
  for (long long w0 = g0; w0 < total; w0 += 2 * stride) {
    uint4 ov[2], dv[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long w = w0 + u * stride;
      ok[u] = w < total;
      if (ok[u]) {
        const int h = (int)(w % H);
        const long long bq = w / H;
        const int q = (int)(bq % Sq), b = (int)(bq / Sq);
        ov[u] = *reinterpret_cast<const uint4*>(o + (long long)b * o_sb + (long long)q * ldo + h * AT + sub * 8);
        dv[u] = *reinterpret_cast<const uint4*>(dout + (long long)b * do_sb + (long long)q * lddo + h * AT + sub * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float s = 0.f;
      if (ok[u]) {
        const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov[u]);
        const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dv[u]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __bfloat1622float2(o2[i]), c = __bfloat1622float2(d2[i]);
          s += a.x * c.x + a.y * c.y;
        }
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      const long long w = w0 + u * stride;
      if (ok[u] && sub == 0) {
        const int h = (int)(w % H);
        const long long bq = w / H;
        dvec[((long long)(bq / Sq) * H + h) * Sq + (bq % Sq)] = s;
      }
    }
  }
}

template <typename K>
static int set_smem(K kern, int bytes, bool& done) {
  if (done) return 0;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return pb_set_cuda_error("cudaFuncSetAttribute(attn)", e);
  done = true;
  return 0;
}

}  // namespace pb

using namespace pb;

static int attn_tmap(CUtensorMap* m, const void* ptr, int S, long long ld, int H, int B, long long sb, int box_rows = AT) {
  return pb_make_tmap_bf16(m, ptr, AT, (uint64_t)S, ld, H, AT, B, sb, 64, (uint32_t)box_rows);
}

static int attn_check(const pb_attn_desc* d) {
  if (d->hd != AT) return pb_set_error("pb_attn: head_dim must be 128 for the tcgen05 attention path");
  if (d->B <= 0 || d->H <= 0 || d->Sq <= 0 || d->Sk <= 0) return pb_set_error("pb_attn: empty problem");
  if (d->causal && d->Sq != d->Sk) return pb_set_error("pb_attn: causal needs Sq == Sk");
  return 0;
}

static int g_attn_dbg_delay = 0;
extern "C" int pb_debug_set_attn_delay(int max_cycles) {
  const int prev = g_attn_dbg_delay;
  g_attn_dbg_delay = max_cycles > 0 ? max_cycles : 0;
  return prev;
}

static void fill_params(AttnParams& p, const pb_attn_desc* d) {
  p.dbg_delay = g_attn_dbg_delay;
  p.B = d->B; p.H = d->H; p.Sq = d->Sq; p.Sk = d->Sk; p.causal = d->causal; p.scale = d->scale;
  p.key_keep = d->key_keep;
  p.o = (__nv_bfloat16*)d->o; p.ldo = d->ldo; p.o_sb = (long long)d->Sq * d->ldo;
  p.lse = d->lse; p.dvec = d->dvec;
  p.dq = (__nv_bfloat16*)d->dq; p.lddq = d->lddq; p.dq_sb = (long long)d->Sq * d->lddq;
  p.dk = (__nv_bfloat16*)d->dk; p.lddk = d->lddk; p.dk_sb = (long long)d->Sk * d->lddk;
  p.dv = (__nv_bfloat16*)d->dv; p.lddv = d->lddv; p.dv_sb = (long long)d->Sk * d->lddv;
}

extern "C" int pb_attn_fwd(const pb_attn_desc* d, void* stream_) {
  if (attn_check(d)) return -1;
  if (d->Sk > FWD_MAX_SK) return pb_set_error("pb_attn_fwd: Sk > 8192 not supported (in-kernel key-padding bitmap)");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CUtensorMap tq, tk, tv, to;
  if (attn_tmap(&tq, d->q, d->Sq, d->ldq, d->H, d->B, (long long)d->Sq * d->ldq)) return -1;
  if (attn_tmap(&tk, d->k, d->Sk, d->ldk, d->H, d->B, (long long)d->Sk * d->ldk)) return -1;
  if (attn_tmap(&tv, d->v, d->Sk, d->ldv, d->H, d->B, (long long)d->Sk * d->ldv)) return -1;
  if (attn_tmap(&to, d->o, d->Sq, d->ldo, d->H, d->B, (long long)d->Sq * d->ldo)) return -1;
  AttnParams p;
  fill_params(p, d);
  dim3 grid((d->Sq + AT - 1) / AT, d->H, d->B);
  // PIANOBART_B200_ATTN_FWD: 3 (default) two Q tiles per CTA; 1 the round-1 one-tile look-ahead kernel (kept as the baseline of
  // profiles/r2_summary.md section 7 and for tools/attn_late_tile_demo.py)
  static const int variant = []() { const char* e = getenv("PIANOBART_B200_ATTN_FWD"); return e ? atoi(e) : 3; }();
  if (variant == 3) {
    static bool attr3 = false;
    const int smem3 = 6 * TILE_BYTES + 1024;
    if (set_smem(attn_fwd3_kernel, smem3, attr3)) return -1;
    dim3 grid3((d->Sq + 2 * AT - 1) / (2 * AT), d->H, d->B);
    PB_LAUNCH(attn_fwd3_kernel, grid3, F3_THREADS, smem3, stream, tq, tk, tv, to, p);
    return pb_check_launch("attn_fwd3_kernel");
  }
  static bool attr = false;
  const int smem = 6 * TILE_BYTES + 1024;
  if (set_smem(attn_fwd_kernel, smem, attr)) return -1;
  PB_LAUNCH(attn_fwd_kernel, grid, NTHREADS, smem, stream, tq, tk, tv, to, p);
  return pb_check_launch("attn_fwd_kernel");
}

// D = rowsum(dO * O) per (batch, head, query): HBM-bound, small footprint - the engine runs it on its side stream next to a
// weight-gradient GEMM (pb_attn_bwd_prep + pb_attn_bwd_main); pb_attn_bwd is the two back to back
extern "C" int pb_attn_bwd_prep(const pb_attn_desc* d, void* stream_) {
  if (attn_check(d)) return -1;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long rows = (long long)d->B * d->Sq * d->H;
  long long blocks = (rows * 16 / 2 + 255) / 256;                       // two rows per thread
  const long long cap = (long long)pb_num_sms() * 16;
  const int grid = (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
  PB_LAUNCH(attn_bwd_prep_kernel, grid, 256, 0, stream, (const __nv_bfloat16*)d->o, (const __nv_bfloat16*)d->dout, d->dvec, d->B, d->H,
            d->Sq, d->ldo, (long long)d->Sq * d->ldo, d->lddo, (long long)d->Sq * d->lddo);
  return pb_check_launch("attn_bwd_prep_kernel");
}

extern "C" int pb_attn_bwd_main(const pb_attn_desc* d, void* stream_) {
  if (attn_check(d)) return -1;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CUtensorMap tq, tk, tv, tdo;
  if (attn_tmap(&tq, d->q, d->Sq, d->ldq, d->H, d->B, (long long)d->Sq * d->ldq)) return -1;
  if (attn_tmap(&tk, d->k, d->Sk, d->ldk, d->H, d->B, (long long)d->Sk * d->ldk)) return -1;
  if (attn_tmap(&tv, d->v, d->Sk, d->ldv, d->H, d->B, (long long)d->Sk * d->ldv)) return -1;
  if (attn_tmap(&tdo, d->dout, d->Sq, d->lddo, d->H, d->B, (long long)d->Sq * d->lddo)) return -1;
  AttnParams p;
  fill_params(p, d);
  static bool attr1 = false, attr2 = false;
  const int smem1 = 2 * TILE_BYTES + 2 * NQ * QT_BYTES + 1024, smem2 = 7 * TILE_BYTES + 1024;
  if (set_smem(attn_bwd_dkv_kernel, smem1, attr1)) return -1;
  if (set_smem(attn_bwd_dq_kernel, smem2, attr2)) return -1;
  CUtensorMap tdq, tdk, tdv;   // output maps (TMA-store epilogues)
  if (attn_tmap(&tdq, d->dq, d->Sq, d->lddq, d->H, d->B, (long long)d->Sq * d->lddq)) return -1;
  if (attn_tmap(&tdk, d->dk, d->Sk, d->lddk, d->H, d->B, (long long)d->Sk * d->lddk)) return -1;
  if (attn_tmap(&tdv, d->dv, d->Sk, d->lddv, d->H, d->B, (long long)d->Sk * d->lddv)) return -1;
  CUtensorMap tq64, tdo64;   // 64-query boxes for the dK/dV kernel
  if (attn_tmap(&tq64, d->q, d->Sq, d->ldq, d->H, d->B, (long long)d->Sq * d->ldq, QB)) return -1;
  if (attn_tmap(&tdo64, d->dout, d->Sq, d->lddo, d->H, d->B, (long long)d->Sq * d->lddo, QB)) return -1;
  dim3 g1((d->Sk + AT - 1) / AT, d->H, d->B);
  PB_LAUNCH(attn_bwd_dkv_kernel, g1, NTHREADS, smem1, stream, tq64, tk, tv, tdo64, tdk, tdv, p);
  if (pb_check_launch("attn_bwd_dkv_kernel")) return -1;
  dim3 g2((d->Sq + AT - 1) / AT, d->H, d->B);
  PB_LAUNCH(attn_bwd_dq_kernel, g2, NTHREADS, smem2, stream, tq, tk, tv, tdo, tdq, p);
  return pb_check_launch("attn_bwd_dq_kernel");
}

extern "C" int pb_attn_bwd(const pb_attn_desc* d, void* stream_) {
  if (pb_attn_bwd_prep(d, stream_)) return -1;
  return pb_attn_bwd_main(d, stream_);
}

#ifdef PB_TRACE
extern "C" int pb_debug_trace(long long* out, int n) {
  if (n > 3 * 64 * 8) n = 3 * 64 * 8;
  return cudaMemcpyFromSymbol(out, pb::pb_trace_buf, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif

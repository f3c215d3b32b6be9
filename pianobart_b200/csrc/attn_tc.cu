// Fused (flash-style) attention for sm_100a: tcgen05 MMAs with S / dP / accumulators in TMEM, operands staged
// by TMA, masks derived in-kernel (key-padding bitmap + causal index compare) - no (B,H,S,S) tensor ever
// reaches HBM.  Replaces HF BartAttention's core (eager_attention_forward: softmax(QK^T*hd^-0.5 + mask) V) and
// its autograd for the three PianoBART variants: encoder self (key padding), decoder self (causal & padding),
// decoder cross (encoder key padding).  head_dim is fixed at 128 (default model).
//
// Three kernels share one tile convention: every operand tile is [128 rows x 128 cols] bf16 stored as two
// 64-column halves of [128 x 128 B] with the 128-byte swizzle.  The same physical tile can be consumed
//   * K-major  : rows = M/N index, columns = contraction index  (Q, K for S=QK^T; dO, V for dP=dO V^T; P, dS as A)
//   * MN-major : rows = contraction index, columns = M/N index  (V for O=PV; P^T, dS^T, dO, Q, K in backward)
// so P / dS are written to shared memory once and feed both dQ = dS K and dK = dS^T Q.
//
//   attn_fwd     CTA = (b, h, 128 queries)  loop over key blocks:  S=QK^T -> online softmax -> O += P V ; LSE out
//   attn_bwd_dkv CTA = (b, h, 128 keys)     loop over query blocks: S, dP -> P, dS -> dV += P^T dO, dK += dS^T Q
//   attn_bwd_dq  CTA = (b, h, 128 queries)  loop over key blocks:  S, dP -> dS -> dQ += dS K
//   attn_bwd_prep                           D = rowsum(dO * O)
// Warp roles per CTA (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax / epilogue
// (thread <-> TMEM lane <-> tile row, so row reductions are thread-local).
#include "ptx.cuh"
#include "pb_internal.h"

namespace pb {

constexpr int AT = 128;                 // tile edge (queries, keys and head_dim)
constexpr int TILE_BYTES = AT * AT * 2;  // 32 KB
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr float LOG2E = 1.4426950408889634f;

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
};

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
__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct Smem4 { uint8_t* t[6]; uint32_t a[6]; };
__device__ __forceinline__ void carve(uint8_t* raw, Smem4& s, int n) {
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  for (int i = 0; i < n; ++i) { s.t[i] = gen + i * TILE_BYTES; s.a[i] = base + i * TILE_BYTES; }
}

// ===================================================================================== forward
__global__ void __launch_bounds__(192, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                const __grid_constant__ CUtensorMap tv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full, v_full, k_empty, v_empty, s_full, p_full, o_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint8_t s_keep[AT];
  Smem4 sm;
  carve(smem_raw, sm, 4);  // 0 Q, 1 K, 2 V, 3 P
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * AT;
  int nkb = (p.Sk + AT - 1) / AT;
  if (p.causal) nkb = min(nkb, qb + 1);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1); mbar_init(&k_full, 1); mbar_init(&v_full, 1); mbar_init(&k_empty, 1); mbar_init(&v_empty, 1);
    mbar_init(&s_full, 1); mbar_init(&p_full, 128); mbar_init(&o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t tS = tmem, tO = tmem + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&q_full, TILE_BYTES);
      load_tile(sm.t[0], &tq, &q_full, q0, h, b);
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(&k_empty, (j & 1) ^ 1);
        mbar_expect_tx(&k_full, TILE_BYTES);
        load_tile(sm.t[1], &tk, &k_full, j * AT, h, b);
        mbar_wait(&v_empty, (j & 1) ^ 1);
        mbar_expect_tx(&v_full, TILE_BYTES);
        load_tile(sm.t[2], &tv, &v_full, j * AT, h, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(&q_full, 0);
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(&k_full, j & 1);
        tc_fence_after();
        mma_tile<false, false>(tS, sm.a[0], sm.a[1], false);      // S = Q K^T
        umma_commit(&s_full);
        umma_commit(&k_empty);
        mbar_wait(&p_full, j & 1);
        mbar_wait(&v_full, j & 1);
        tc_fence_after();
        mma_tile<false, true>(tO, sm.a[3], sm.a[2], false);       // O_part = P V   (V as MN-major B)
        umma_commit(&o_full);
        umma_commit(&v_empty);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;          // tile row = TMEM lane
    const int tid = threadIdx.x - 64;        // 0..127
    const int qg = q0 + r;                   // global query index
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    float o[AT];
#pragma unroll
    for (int i = 0; i < AT; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nkb; ++j) {
      const int kg0 = j * AT;
      {
        const int kc = kg0 + tid;
        uint8_t kp = (kc < p.Sk) ? 1 : 0;
        if (kp && p.key_keep) kp = p.key_keep[(long long)b * p.Sk + kc];
        compute_bar_sync();                  // previous iteration finished reading s_keep
        s_keep[tid] = kp;
        compute_bar_sync();
      }
      mbar_wait(&s_full, j & 1);
      tc_fence_after();
      // pass 1: masked row maximum
      float bm = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i;
          const bool ok = s_keep[col] && (!p.causal || kg0 + col <= qg);
          if (ok) bm = fmaxf(bm, __uint_as_float(v[i]) * sl2);
        }
      }
      const float m_new = fmaxf(m, bm);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m == -INFINITY) ? 0.f : exp2f(m - m_use);
      float rs = 0.f;
      // pass 2: probabilities -> bf16 -> swizzled smem tile (A operand of the PV product)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_addr + c * 32, v);
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i;
          const bool ok = s_keep[col] && (!p.causal || kg0 + col <= qg);
          const float e = ok ? exp2f(__uint_as_float(v[i]) * sl2 - m_use) : 0.f;
          // the row sum must match what the tensor core will see: accumulate the bf16-rounded value
          x[i] = __bfloat162float(__float2bfloat16(e));
          rs += x[i];
        }
        store_chunk(sm.t[3], r, c * 32, x);
      }
      l = l * alpha + rs;
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full);
      mbar_wait(&o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tO + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
    }
    if (qg < p.Sq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* orow = p.o + (long long)b * p.o_sb + (long long)qg * p.ldo + h * AT;
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        uint4 v;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(o[g * 8 + 2 * t] * inv, o[g * 8 + 2 * t + 1] * inv);
        *reinterpret_cast<uint4*>(orow + g * 8) = v;
      }
      p.lse[((long long)b * p.H + h) * p.Sq + qg] = (l > 0.f) ? (m + log2f(l)) : INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 256);
}

// ===================================================================================== backward: dK, dV
__global__ void __launch_bounds__(192, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                    const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t kv_full, qdo_full, qdo_empty, sdp_full, pds_full, acc_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint8_t s_keep[AT];
  Smem4 sm;
  carve(smem_raw, sm, 6);  // 0 K, 1 V, 2 Q, 3 dO, 4 P, 5 dS
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kb * AT;
  const int nqb = (p.Sq + AT - 1) / AT;
  const int qb0 = p.causal ? kb : 0;       // causal: only query blocks at or below the diagonal see these keys
  const int niter = nqb - qb0;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo); }
  if (warp == 1 && lane == 0) {
    mbar_init(&kv_full, 1); mbar_init(&qdo_full, 1); mbar_init(&qdo_empty, 1); mbar_init(&sdp_full, 1);
    mbar_init(&pds_full, 128); mbar_init(&acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 384;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&kv_full, 2 * TILE_BYTES);
      load_tile(sm.t[0], &tk, &kv_full, k0, h, b);
      load_tile(sm.t[1], &tv, &kv_full, k0, h, b);
      for (int it = 0; it < niter; ++it) {
        mbar_wait(&qdo_empty, (it & 1) ^ 1);
        mbar_expect_tx(&qdo_full, 2 * TILE_BYTES);
        load_tile(sm.t[2], &tq, &qdo_full, (qb0 + it) * AT, h, b);
        load_tile(sm.t[3], &tdo, &qdo_full, (qb0 + it) * AT, h, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(&kv_full, 0);
      for (int it = 0; it < niter; ++it) {
        mbar_wait(&qdo_full, it & 1);
        tc_fence_after();
        mma_tile<false, false>(tS, sm.a[2], sm.a[0], false);    // S  = Q K^T
        mma_tile<false, false>(tdP, sm.a[3], sm.a[1], false);   // dP = dO V^T
        umma_commit(&sdp_full);
        mbar_wait(&pds_full, it & 1);
        tc_fence_after();
        mma_tile<true, true>(tdV, sm.a[4], sm.a[3], it > 0);    // dV += P^T dO
        mma_tile<true, true>(tdK, sm.a[5], sm.a[2], it > 0);    // dK += dS^T Q
        umma_commit(&qdo_empty);                                 // Q, dO, P, dS buffers reusable
      }
      umma_commit(&acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    {
      const int kc = k0 + tid;
      uint8_t kp = (kc < p.Sk) ? 1 : 0;
      if (kp && p.key_keep) kp = p.key_keep[(long long)b * p.Sk + kc];
      s_keep[tid] = kp;
      compute_bar_sync();
    }
    const long long rbase = ((long long)b * p.H + h) * p.Sq;
    float L_next = INFINITY, D_next = 0.f;
    if (qb0 * AT + r < p.Sq) { L_next = p.lse[rbase + qb0 * AT + r]; D_next = p.dvec[rbase + qb0 * AT + r]; }
    for (int it = 0; it < niter; ++it) {
      const int qg = (qb0 + it) * AT + r;
      const bool qok = qg < p.Sq;
      const float L = L_next, Dv = D_next;
      if (it + 1 < niter && qg + AT < p.Sq) { L_next = p.lse[rbase + qg + AT]; D_next = p.dvec[rbase + qg + AT]; }
      else { L_next = INFINITY; D_next = 0.f; }
      mbar_wait(&sdp_full, it & 1);
      tc_fence_after();
      // P / dS smem of the previous iteration were released by qdo_empty's MMAs; the MMA warp only issues this
      // iteration's S/dP after them (in-order tensor pipe), so sdp_full implies the buffers are free.
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld32(tS + lane_addr + c * 32, sv);
        tmem_ld32(tdP + lane_addr + c * 32, dv);
        tmem_ld_wait();
        float pr[32], ds[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i;
          const bool ok = qok && s_keep[col] && (!p.causal || k0 + col <= qg);
          const float pv = ok ? exp2f(__uint_as_float(sv[i]) * sl2 - L) : 0.f;
          pr[i] = pv;
          ds[i] = pv * (__uint_as_float(dv[i]) - Dv) * p.scale;
        }
        store_chunk(sm.t[4], r, c * 32, pr);
        store_chunk(sm.t[5], r, c * 32, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&pds_full);
    }
    // epilogue: thread = key row
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int kg = k0 + r;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t tacc = which == 0 ? tdV : tdK;
      __nv_bfloat16* dst = which == 0 ? (p.dv + (long long)b * p.dv_sb + (long long)kg * p.lddv + h * AT)
                                      : (p.dk + (long long)b * p.dk_sb + (long long)kg * p.lddk + h * AT);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tacc + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (kg < p.Sk) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              h2[t] = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * t]), __uint_as_float(v[g * 8 + 2 * t + 1]));
            *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = o;
          }
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ===================================================================================== backward: dQ
__global__ void __launch_bounds__(192, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t qdo_full, kv_full, kv_empty, sdp_full, ds_full, acc_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint8_t s_keep[AT];
  Smem4 sm;
  carve(smem_raw, sm, 5);  // 0 Q, 1 dO, 2 K, 3 V, 4 dS
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * AT;
  int nkb = (p.Sk + AT - 1) / AT;
  if (p.causal) nkb = min(nkb, qb + 1);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo); }
  if (warp == 1 && lane == 0) {
    mbar_init(&qdo_full, 1); mbar_init(&kv_full, 1); mbar_init(&kv_empty, 1); mbar_init(&sdp_full, 1);
    mbar_init(&ds_full, 128); mbar_init(&acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&qdo_full, 2 * TILE_BYTES);
      load_tile(sm.t[0], &tq, &qdo_full, q0, h, b);
      load_tile(sm.t[1], &tdo, &qdo_full, q0, h, b);
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(&kv_empty, (j & 1) ^ 1);
        mbar_expect_tx(&kv_full, 2 * TILE_BYTES);
        load_tile(sm.t[2], &tk, &kv_full, j * AT, h, b);
        load_tile(sm.t[3], &tv, &kv_full, j * AT, h, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(&qdo_full, 0);
      for (int j = 0; j < nkb; ++j) {
        mbar_wait(&kv_full, j & 1);
        tc_fence_after();
        mma_tile<false, false>(tS, sm.a[0], sm.a[2], false);    // S  = Q K^T
        mma_tile<false, false>(tdP, sm.a[1], sm.a[3], false);   // dP = dO V^T
        umma_commit(&sdp_full);
        mbar_wait(&ds_full, j & 1);
        tc_fence_after();
        mma_tile<false, true>(tdQ, sm.a[4], sm.a[2], j > 0);    // dQ += dS K   (K as MN-major B)
        umma_commit(&kv_empty);
      }
      umma_commit(&acc_full);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int tid = threadIdx.x - 64;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const int qg = q0 + r;
    const bool qok = qg < p.Sq;
    const long long ridx = ((long long)b * p.H + h) * p.Sq + qg;
    const float L = qok ? p.lse[ridx] : INFINITY;
    const float Dv = qok ? p.dvec[ridx] : 0.f;
    for (int j = 0; j < nkb; ++j) {
      const int kg0 = j * AT;
      {
        const int kc = kg0 + tid;
        uint8_t kp = (kc < p.Sk) ? 1 : 0;
        if (kp && p.key_keep) kp = p.key_keep[(long long)b * p.Sk + kc];
        compute_bar_sync();
        s_keep[tid] = kp;
        compute_bar_sync();
      }
      mbar_wait(&sdp_full, j & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld32(tS + lane_addr + c * 32, sv);
        tmem_ld32(tdP + lane_addr + c * 32, dv);
        tmem_ld_wait();
        float ds[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i;
          const bool ok = qok && s_keep[col] && (!p.causal || kg0 + col <= qg);
          const float pv = ok ? exp2f(__uint_as_float(sv[i]) * sl2 - L) : 0.f;
          ds[i] = pv * (__uint_as_float(dv[i]) - Dv) * p.scale;
        }
        store_chunk(sm.t[4], r, c * 32, ds);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&ds_full);
    }
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    __nv_bfloat16* dst = p.dq + (long long)b * p.dq_sb + (long long)qg * p.lddq + h * AT;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(tdQ + lane_addr + c * 32, v);
      tmem_ld_wait();
      if (qok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            h2[t] = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * t]), __uint_as_float(v[g * 8 + 2 * t + 1]));
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = o;
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// D[b,h,q] = sum_c dO[b,q,h,c] * O[b,q,h,c]      (one warp per (b,q,h) row of 128)
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                                            float* __restrict__ dvec, int B, int H, int Sq, long long ldo,
                                                            long long o_sb, long long lddo, long long do_sb) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * Sq * H;
  if (w >= total) return;
  const int h = (int)(w % H);
  const long long bq = w / H;
  const int q = (int)(bq % Sq), b = (int)(bq / Sq);
  const uint2 ov = *reinterpret_cast<const uint2*>(o + (long long)b * o_sb + (long long)q * ldo + h * AT + lane * 4);
  const uint2 dv = *reinterpret_cast<const uint2*>(dout + (long long)b * do_sb + (long long)q * lddo + h * AT + lane * 4);
  const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
  const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dv);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float2 a = __bfloat1622float2(o2[i]), c = __bfloat1622float2(d2[i]);
    s += a.x * c.x + a.y * c.y;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) dvec[((long long)b * H + h) * Sq + q] = s;
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

static int attn_tmap(CUtensorMap* m, const void* ptr, int S, long long ld, int H, int B, long long sb) {
  return pb_make_tmap_bf16(m, ptr, AT, (uint64_t)S, ld, H, AT, B, sb, 64, AT);
}

static int attn_check(const pb_attn_desc* d) {
  if (d->hd != AT) return pb_set_error("pb_attn: head_dim must be 128 for the tcgen05 attention path");
  if (d->B <= 0 || d->H <= 0 || d->Sq <= 0 || d->Sk <= 0) return pb_set_error("pb_attn: empty problem");
  if (d->causal && d->Sq != d->Sk) return pb_set_error("pb_attn: causal needs Sq == Sk");
  return 0;
}

static void fill_params(AttnParams& p, const pb_attn_desc* d) {
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
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CUtensorMap tq, tk, tv;
  if (attn_tmap(&tq, d->q, d->Sq, d->ldq, d->H, d->B, (long long)d->Sq * d->ldq)) return -1;
  if (attn_tmap(&tk, d->k, d->Sk, d->ldk, d->H, d->B, (long long)d->Sk * d->ldk)) return -1;
  if (attn_tmap(&tv, d->v, d->Sk, d->ldv, d->H, d->B, (long long)d->Sk * d->ldv)) return -1;
  AttnParams p;
  fill_params(p, d);
  static bool attr = false;
  const int smem = 4 * TILE_BYTES + 1024;
  if (set_smem(attn_fwd_kernel, smem, attr)) return -1;
  dim3 grid((d->Sq + AT - 1) / AT, d->H, d->B);
  attn_fwd_kernel<<<grid, 192, smem, stream>>>(tq, tk, tv, p);
  return pb_check_launch("attn_fwd_kernel");
}

extern "C" int pb_attn_bwd(const pb_attn_desc* d, void* stream_) {
  if (attn_check(d)) return -1;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CUtensorMap tq, tk, tv, tdo;
  if (attn_tmap(&tq, d->q, d->Sq, d->ldq, d->H, d->B, (long long)d->Sq * d->ldq)) return -1;
  if (attn_tmap(&tk, d->k, d->Sk, d->ldk, d->H, d->B, (long long)d->Sk * d->ldk)) return -1;
  if (attn_tmap(&tv, d->v, d->Sk, d->ldv, d->H, d->B, (long long)d->Sk * d->ldv)) return -1;
  if (attn_tmap(&tdo, d->dout, d->Sq, d->lddo, d->H, d->B, (long long)d->Sq * d->lddo)) return -1;
  AttnParams p;
  fill_params(p, d);
  {
    const long long rows = (long long)d->B * d->Sq * d->H;
    const int grid = (int)((rows * 32 + 255) / 256);
    attn_bwd_prep_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)d->o, (const __nv_bfloat16*)d->dout, d->dvec, d->B, d->H,
                                                   d->Sq, d->ldo, (long long)d->Sq * d->ldo, d->lddo, (long long)d->Sq * d->lddo);
    if (pb_check_launch("attn_bwd_prep_kernel")) return -1;
  }
  static bool attr1 = false, attr2 = false;
  const int smem1 = 6 * TILE_BYTES + 1024, smem2 = 5 * TILE_BYTES + 1024;
  if (set_smem(attn_bwd_dkv_kernel, smem1, attr1)) return -1;
  if (set_smem(attn_bwd_dq_kernel, smem2, attr2)) return -1;
  dim3 g1((d->Sk + AT - 1) / AT, d->H, d->B);
  attn_bwd_dkv_kernel<<<g1, 192, smem1, stream>>>(tq, tk, tv, tdo, p);
  if (pb_check_launch("attn_bwd_dkv_kernel")) return -1;
  dim3 g2((d->Sq + AT - 1) / AT, d->H, d->B);
  attn_bwd_dq_kernel<<<g2, 192, smem2, stream>>>(tq, tk, tv, tdo, p);
  return pb_check_launch("attn_bwd_dq_kernel");
}

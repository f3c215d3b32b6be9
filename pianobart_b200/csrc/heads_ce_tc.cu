// Fused MLM heads + masked cross-entropy for sm_100a (north-star fusion 3): the eight Linear heads of model.py:119-126
// (one [V = 1280, d] weight), the per-attribute masked CE of pretrain.py:112-118 with the mis-ordered weights of :179-189,
// the argmax accuracy of :163-176 and the gradient wrt the logits, in ONE kernel - the fp32 logits [M, V] (84 MB per step at
// the default batch) never reach HBM; only bf16 dlogits are written for the two backward GEMMs.
//
// CTA = 128 rows of the decoder output h.  The vocabulary is cut into column GROUPS of whole attribute segments that fit
// the 512 TMEM columns (default Octuple sizes: {Bar, Position} | {Instrument, Pitch} | {Duration, Velocity, TimeSig, Tempo}
// = 416 / 416 / 496 accumulator columns).  Per group:
//   warp 0  TMA producer: 16 k-blocks of [128 x 64] h and [N_g x 64] W tiles through a 2-stage ring
//   warp 1  MMA issuer:   per k16 step two tcgen05.mma (N = 256 and N_g - 256), accumulating logits in TMEM
//   warps 2-5 epilogue, thread = row: per segment  pass 1 max / argmax / target logit (bias added on the fly),
//           pass 2 e = exp2((x - max) log2 e) written back over x in TMEM + row sum; then one output pass over the group's
//           columns: dlogit = (e / sum - onehot) * mask * w / (sum_w * den) -> bf16, transposed through a per-warp
//           shared-memory tile so that global stores are 64-byte row pieces.
// A group starts at a multiple of 8 columns (16-byte aligned stores), so its first few columns may belong to the previous
// group's last segment: the thread still holds that segment's statistics (shared-memory stats row) and finishes them here.
// The producer prefetches the next group's first tiles while the epilogue drains TMEM.
#include "ptx.cuh"
#include "pb_internal.h"

namespace pb {

constexpr int HM = 128;                       // rows per CTA
constexpr int HBK = 64;                       // k-block (one 128-byte swizzle span of bf16)
constexpr int HSTAGES = 2;
constexpr int HNMAX = 496;                    // accumulator columns of a group (<= 512 TMEM columns)
constexpr int HA_BYTES = HM * HBK * 2;        // 16 KB
constexpr int HB_BYTES = HNMAX * HBK * 2;     // 62 KB
constexpr int HSTAGE_BYTES = HA_BYTES + HB_BYTES;
constexpr int HDYN_BYTES = HSTAGES * HSTAGE_BYTES + 1024;
constexpr int HMAXG = 4, HMAXSEG = 8, HMAXV = 1536;
constexpr int HTHREADS = 64 + 128;
constexpr float H_LOG2E = 1.4426950408889634f, H_LN2 = 0.6931471805599453f;

struct HeadsFusedParams {
  const float* bias; const int* targets; const float* mask; const float* den;
  float* loss_num; float* correct; __nv_bfloat16* dlogits; int* argmax_out;
  long long M;
  int K, V, nseg, ngroups;
  int off[HMAXSEG + 1];
  float w[HMAXSEG];
  float sum_w, grad_scale;
  int g_origin[HMAXG], g_n[HMAXG], g_first[HMAXG], g_last[HMAXG], g_wend[HMAXG];
};

__device__ __forceinline__ float h_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float h_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(HTHREADS, 1)
heads_ce_fused_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb0,
                      const __grid_constant__ CUtensorMap tb1, const __grid_constant__ CUtensorMap tb2,
                      const __grid_constant__ CUtensorMap tb3, const __grid_constant__ HeadsFusedParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[HSTAGES], empty_bar[HSTAGES], acc_full, acc_empty, mma_drain;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_bias[HMAXV];
  __shared__ __align__(16) float4 s_stat[HM][HMAXSEG + 1];     // {max, 1/sum, coef, target column (int bits)}; 9 float4 per row: conflict-free
  __shared__ __align__(16) uint8_t s_out[4][32][80];           // per-warp transpose tile: 32 rows x (64 B + 16 B pad)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * HM;
  const int nkb = P.K / HBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&ta); tma_prefetch_desc(&tb0); tma_prefetch_desc(&tb1); tma_prefetch_desc(&tb2); tma_prefetch_desc(&tb3);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < HSTAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&acc_full, 1); mbar_init(&acc_empty, 128); mbar_init(&mma_drain, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = tmem_base_smem;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int g = 0; g < P.ngroups; ++g) {
        const CUtensorMap* tb = g == 0 ? &tb0 : (g == 1 ? &tb1 : (g == 2 ? &tb2 : &tb3));
        const int ng = P.g_n[g], origin = P.g_origin[g];
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % HSTAGES;
          mbar_wait(&empty_bar[s], ((uint32_t)(it / HSTAGES) & 1) ^ 1);
          uint8_t* a_dst = gen + s * HSTAGE_BYTES;
          uint8_t* b_dst = a_dst + HA_BYTES;
          mbar_expect_tx(&full_bar[s], HA_BYTES + ng * HBK * 2);
          tma_load_4d(a_dst, &ta, &full_bar[s], kb * HBK, (int)m0, 0, 0);
          tma_load_4d(b_dst, tb, &full_bar[s], kb * HBK, origin, 0, 0);
          tma_load_4d(b_dst + (ng / 2) * HBK * 2, tb, &full_bar[s], kb * HBK, origin + ng / 2, 0, 0);
        }
      }
      // producer tail: the last slot releases (tcgen05.commit arrivals) have landed before the CTA exits
      for (int i = it; i < it + HSTAGES; ++i) mbar_wait(&empty_bar[i % HSTAGES], ((uint32_t)(i / HSTAGES) & 1) ^ 1);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int it = 0;
      for (int g = 0; g < P.ngroups; ++g) {
        const int ng = P.g_n[g];
        const int n1 = ng < 256 ? ng : 256, n2 = ng - n1;
        const uint32_t idesc1 = make_idesc_bf16(HM, n1, 0, 0), idesc2 = make_idesc_bf16(HM, n2 > 0 ? n2 : 16, 0, 0);
        if (g > 0) {                         // the epilogue has drained the previous group's accumulator
          mbar_wait(&acc_empty, (uint32_t)(g - 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % HSTAGES;
          mbar_wait(&full_bar[s], (uint32_t)(it / HSTAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = base + s * HSTAGE_BYTES, b_addr = a_addr + HA_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + kk * 32, 16, 1024);
            const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
            umma_bf16(tmem, ad, make_smem_desc_sw128(b_addr + kk * 32, 16, 1024), idesc1, acc);
            if (n2 > 0) umma_bf16(tmem + 256, ad, make_smem_desc_sw128(b_addr + 256 * 128 + kk * 32, 16, 1024), idesc2, acc);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full);
      }
      umma_commit(&mma_drain);
      mbar_wait(&mma_drain, 0);
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = row
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const long long row = m0 + r;
    const bool row_ok = row < P.M;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16);
    uint8_t (*so)[80] = s_out[quad];
    for (int i = threadIdx.x - 64; i < P.V; i += 128) s_bias[i] = P.bias[i];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float loss_acc[HMAXSEG], cor_acc[HMAXSEG];
#pragma unroll
    for (int s = 0; s < HMAXSEG; ++s) { loss_acc[s] = 0.f; cor_acc[s] = 0.f; }

    // e = exp2((x + bias - mx) log2 e) over relative columns [lo, hi) of the current group, written back over x; returns the sum
    auto exp_range = [&](int origin, int lo, int hi, float mx) -> float {
      float sum0 = 0.f, sum1 = 0.f;
      const float nm = -mx * H_LOG2E;
      for (int c32 = lo & ~31; c32 < hi; c32 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c32, v);
        tmem_ld_wait();
        const bool full = c32 >= lo && c32 + 32 <= hi;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int c = c32 + i;
          const float e0 = h_ex2(fmaf(__uint_as_float(v[i]) + s_bias[origin + c], H_LOG2E, nm));
          const float e1 = h_ex2(fmaf(__uint_as_float(v[i + 1]) + s_bias[origin + c + 1], H_LOG2E, nm));
          if (full || (c >= lo && c < hi)) { v[i] = __float_as_uint(e0); sum0 += e0; }
          if (full || (c + 1 >= lo && c + 1 < hi)) { v[i + 1] = __float_as_uint(e1); sum1 += e1; }
        }
        tmem_st32(taddr + c32, v);
      }
      tmem_st_wait();
      return sum0 + sum1;
    };

    for (int g = 0; g < P.ngroups; ++g) {
      const int origin = P.g_origin[g], first = P.g_first[g], last = P.g_last[g], wend = P.g_wend[g];
      mbar_wait(&acc_full, (uint32_t)g & 1);
      tc_fence_after();
      // columns in front of the group's first segment belong to the previous segment (statistics already known)
      if (P.off[first] > origin) exp_range(origin, 0, P.off[first] - origin, s_stat[r][first - 1].x);
      for (int s = first; s <= last; ++s) {
        const int lo = P.off[s] - origin, hi = P.off[s + 1] - origin, n = hi - lo;
        const int t = row_ok ? P.targets[row * P.nseg + s] : -1;
        const float mk = row_ok ? P.mask[row * P.nseg + s] : 0.f;
        // pass 1: maximum, argmax (lowest index on ties, pretrain.py:165 np.argmax), target logit
        float mx = -INFINITY, xt = 0.f;
        int am = 0;
        for (int c32 = lo & ~31; c32 < hi; c32 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c32, v);
          tmem_ld_wait();
          const bool full = c32 >= lo && c32 + 32 <= hi;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = c32 + i;
            if (full || (c >= lo && c < hi)) {
              const float x = __uint_as_float(v[i]) + s_bias[origin + c];
              if (x > mx) { mx = x; am = c - lo; }
              if (c - lo == t) xt = x;
            }
          }
        }
        // pass 2: exponentials in place + row sum
        const float sum = exp_range(origin, lo, hi, mx);
        const bool t_ok = t >= 0 && t < n;
        const float coef = (mk != 0.f) ? mk * P.w[s] / (P.sum_w * P.den[s]) * P.grad_scale : 0.f;
        s_stat[r][s] = make_float4(mx, 1.f / sum, coef, __int_as_float(t_ok ? t : -1));
        if (mk != 0.f && t_ok) loss_acc[s] += (mx + log2f(sum) * H_LN2 - xt) * mk;
        if (t_ok && am == t) cor_acc[s] += mk;
        if (P.argmax_out && row_ok) P.argmax_out[row * P.nseg + s] = am;
      }
      // output pass: this group owns the absolute columns [origin, wend)
      int cur = P.off[first] > origin ? first - 1 : first;
      for (int c32 = 0; P.dlogits != nullptr && origin + c32 < wend; c32 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c32, v);
        tmem_ld_wait();
        const int abs0 = origin + c32;
        while (cur < P.nseg - 1 && abs0 >= P.off[cur + 1]) ++cur;
        const int bnd = P.off[cur + 1];                    // columns >= bnd of this chunk belong to segment cur + 1
        const float4 sa = s_stat[r][cur];
        const float4 sb = (cur + 1 < P.nseg && bnd < abs0 + 32) ? s_stat[r][cur + 1] : sa;
        const int ta_col = P.off[cur] + __float_as_int(sa.w), tb_col = bnd + __float_as_int(sb.w);
        const bool ta_ok = __float_as_int(sa.w) >= 0, tb_ok = __float_as_int(sb.w) >= 0;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float gv[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int a = abs0 + i + j;
            const bool second = a >= bnd;
            const float inv = second ? sb.y : sa.y, cf = second ? sb.z : sa.z;
            const bool hot = second ? (tb_ok && a == tb_col) : (ta_ok && a == ta_col);
            gv[j] = cf != 0.f ? (__uint_as_float(v[i + j]) * inv - (hot ? 1.f : 0.f)) * cf : 0.f;
          }
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(gv[0], gv[1]);
          pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(&so[lane][q * 16]) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        __syncwarp();
#pragma unroll
        for (int itr = 0; itr < 4; ++itr) {
          const int rr = itr * 8 + (lane >> 2), piece = lane & 3;
          const long long grow = m0 + quad * 32 + rr;
          const int acol = abs0 + piece * 8;
          if (grow < P.M && acol < wend)
            *reinterpret_cast<uint4*>(P.dlogits + grow * P.V + acol) = *reinterpret_cast<const uint4*>(&so[rr][piece * 16]);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty);
    }
#pragma unroll
    for (int s = 0; s < HMAXSEG; ++s) {
      if (s < P.nseg) {
        const float ls = h_warp_sum(loss_acc[s]), cs = h_warp_sum(cor_acc[s]);
        if (lane == 0) {
          if (ls != 0.f) atomicAdd(P.loss_num + s, ls);
          if (cs != 0.f) atomicAdd(P.correct + s, cs);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace pb

extern "C" int pb_heads_ce_fused(const void* h, long long ldh, const void* w, const float* bias, const int* targets,
                                 const float* mask, const float* den, float* loss_num, float* correct, void* dlogits,
                                 int* argmax_out, long long M, int K, int nseg, const int* seg_sizes_host,
                                 const float* weights_host, float grad_scale, void* stream_) {
  using namespace pb;
  if (nseg < 1 || nseg > HMAXSEG) return pb_set_error("heads_ce_fused: nseg must be in [1,8]");
  if (K < HBK || K % HBK != 0) return pb_set_error("heads_ce_fused: K must be a multiple of 64");
  if (M <= 0) return pb_set_error("heads_ce_fused: empty problem");
  HeadsFusedParams P;
  memset(&P, 0, sizeof(P));
  int off = 0;
  float sw = 0.f;
  for (int s = 0; s < nseg; ++s) {
    if (seg_sizes_host[s] < 32) return pb_set_error("heads_ce_fused: segments must have >= 32 classes");
    P.off[s] = off; off += seg_sizes_host[s]; P.w[s] = weights_host[s]; sw += weights_host[s];
  }
  P.off[nseg] = off;
  if (off % 8 != 0 || off > HMAXV) return pb_set_error("heads_ce_fused: vocabulary must be a multiple of 8 and <= 1536");
  P.V = off; P.K = K; P.nseg = nseg; P.M = M; P.sum_w = sw; P.grad_scale = grad_scale;
  P.bias = bias; P.targets = targets; P.mask = mask; P.den = den; P.loss_num = loss_num; P.correct = correct;
  P.dlogits = reinterpret_cast<__nv_bfloat16*>(dlogits); P.argmax_out = argmax_out;
  // greedy grouping of whole segments into <= 496 accumulator columns, every group starting at a multiple of 8
  int ng = 0, s = 0;
  while (s < nseg) {
    if (ng == HMAXG) return pb_set_error("heads_ce_fused: vocabulary needs more than 4 column groups");
    const int origin = P.off[s] & ~7;
    int last = s;
    if (((P.off[s + 1] - origin + 15) & ~15) > HNMAX) return pb_set_error("heads_ce_fused: segment wider than 496 classes");
    while (last + 1 < nseg && ((P.off[last + 2] - origin + 15) & ~15) <= HNMAX) ++last;
    int n = (P.off[last + 1] - origin + 15) & ~15;
    if (n < 32) n = 32;
    P.g_origin[ng] = origin; P.g_n[ng] = n; P.g_first[ng] = s; P.g_last[ng] = last;
    s = last + 1;
    P.g_wend[ng] = s < nseg ? (P.off[s] & ~7) : P.V;
    ++ng;
  }
  P.ngroups = ng;
  CUtensorMap ta, tb[HMAXG];
  if (pb_make_tmap_bf16(&ta, h, (uint64_t)K, (uint64_t)M, ldh, 1, 0, 1, 0, HBK, HM)) return -1;
  for (int g = 0; g < HMAXG; ++g) {
    const int gg = g < ng ? g : 0;
    if (pb_make_tmap_bf16(&tb[g], w, (uint64_t)K, (uint64_t)P.V, K, 1, 0, 1, 0, HBK, (uint32_t)(P.g_n[gg] / 2))) return -1;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(heads_ce_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HDYN_BYTES);
    if (e != cudaSuccess) return pb_set_cuda_error("cudaFuncSetAttribute(heads_ce_fused)", e);
    attr = true;
  }
  const unsigned grid = (unsigned)((M + HM - 1) / HM);
  PB_LAUNCH(heads_ce_fused_kernel, grid, HTHREADS, HDYN_BYTES, reinterpret_cast<cudaStream_t>(stream_), ta, tb[0], tb[1], tb[2], tb[3], P);
  return pb_check_launch("heads_ce_fused");
}

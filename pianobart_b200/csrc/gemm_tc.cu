// tcgen05 / TMEM / TMA GEMM for sm_100a.
//
//   C[b,h][m,n] = epi( alpha * sum_k A[b,h][m,k] * B[b,h][n,k] )        bf16 x bf16 -> fp32 (TMEM)
//
// Every dense contraction of the PianoBART hot path goes through this kernel:
// the in_linear of the Octuple front end (reference PianoBart.py:68,71), the q/k/v/out
// projections and fc1/fc2 of every BART layer (HF modeling_bart.py BartAttention / Bart*Layer),
// the eight LM heads concatenated to N=1280 (reference model.py:119-126), the batched
// QK^T / PV attention products, and all of their backward products (dX, dW).
//
// Design (B200-first, not a translation of anything in the reference, which has no kernels):
//   * persistent grid, one CTA per SM, 320 threads = 10 warps with fixed roles:
//       warp 0   TMA producer   (cp.async.bulk.tensor 4D, 128B swizzle, mbarrier complete_tx)
//       warp 1   MMA issuer     (one elected thread issues tcgen05.mma, commits to mbarriers)
//       warp 2-9 epilogue       (tcgen05.ld TMEM->registers, bias / GELU + gelu' / dropout / residual; a thread owns
//                                one accumulator row, so C, the residual / aux input and the aux output move as
//                                [32 x 32] boxes staged in shared memory by bulk tensor copies instead of row-per-
//                                thread global accesses; 8 warps = lane quadrant x column half so two warps share
//                                a scheduler; bias staged in smem once per tile; a specialised instantiation (FAST)
//                                carries only the bf16 TMA path the training step uses)
//   * cta_group::2: for M >= 1024 a CTA pair (cluster of 2) computes a 256 x 256 tile, each CTA staging its 128 rows
//     of A and half of B.
//   * CTA tile 128 x BLOCK_N (128 or 256), BLOCK_K = 64 bf16 = one 128-byte swizzle row,
//     multi-stage smem ring (full/empty mbarriers), two TMEM accumulators (2 x BLOCK_N columns)
//     so the epilogue of tile i overlaps the main loop of tile i+1.
//   * both operands may be K-major (row = m or n, k contiguous) or MN-major (row = k, m/n
//     contiguous) so that forward (X W^T), dX (dY W) and dW (dY^T X) all read the tensors in the
//     layout they already have in HBM - no transposed copies.
//   * 4-D tensor maps (inner, rows, h, b) give batched / strided operands (per-head attention
//     slices of the fused QKV activation) without any gather kernel.
//   * split-K with fp32 red.global.add for the dW products whose output tile count cannot fill
//     148 SMs.
#include "ptx.cuh"
#include "pb_internal.h"
#include "dropout.cuh"

#include <mutex>
#include <unordered_map>
#include <string>
#include <cstring>
#include <cstdlib>

namespace pb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;          // (TMEM lane quadrant) x (column half): two warps per scheduler
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;

struct GemmKParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks, kb_per_split, split_k;
  int batch_h, batch_b;
  long long total_units;
  void* c;
  long long ldc, c_stride_h, c_stride_b;
  const float* bias;
  const void* residual;
  long long ldr, r_stride_h, r_stride_b;
  float alpha;
  int flags;
  void* aux;
  long long ldaux;
  int r_row_mod;
  pbdrop::Site drop;
  int causal;  // 1: skip tiles entirely above the diagonal (n0 > m0 + BLOCK_M - 1); 2: limit k range to m0+BLOCK_M
  // epilogue through TMA (bf16, 32 x 32 boxes staged in shared memory, 64-byte swizzle): see the kernel's epilogue
  int tma_c;    // C is stored with tmap_c
  int tma_pre;  // 0: none, 1: residual rows, 2: aux rows (MUL_AUX / MUL_DGELU operand) are loaded with tmap_pre
  int tma_x;    // aux output (AUX_PREACT / AUX_DGELU) is stored with tmap_x
  // Tail splitting (cta_group::2, 256-wide tiles, no batching / split-K): the persistent grid has `slots` CTA pairs; when
  // the tiles left over after the last full wave fill at most half of the slots, each of them is issued as two 256 x 128
  // units (UMMA N = 128 from the same shared-memory stages), so the partial wave costs half a tile time instead of a whole
  // one: 64 x 4 tiles on 74 pairs run in 3.5 instead of 4 tile times (out_proj, dO, q_c, dH*: profiles/r2_summary.md).
  int full_units;   // units [0, full_units) are whole tiles; later units are (tile, column half) pairs.  -1: no splitting
};

// CG2: cta_group::2 - a pair of CTAs (one TPC) computes a 256 x BLOCK_N tile; each CTA stages its own 128 rows of A and
// HALF of the B tile, so per SM the shared-memory traffic (TMA writes + MMA operand reads) drops from ~192 to ~128 B/clk,
// which is what lifts the kernel off the shared-memory roof.
template <int BLOCK_N, bool CG2 = false>
struct SmemCfg {
  static constexpr int B_ROWS = CG2 ? BLOCK_N / 2 : BLOCK_N;   // rows of the B tile staged by one CTA
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue staging per epilogue warp: two 2 KB slots ([32 rows x 32 bf16] boxes) - both rotate as TMA store buffers,
  // or one takes the TMA loads when the epilogue has a residual / aux input.  Sized so that the 32 KB-per-stage
  // configurations keep a 6-deep operand ring (5 stages cost the long-K GEMMs ~5%).
  static constexpr int EPI_WARP_BYTES = 2 * 2048;
  static constexpr int EPI_BYTES = 8 * EPI_WARP_BYTES;
  static constexpr int RING_BUDGET = 227 * 1024 - 1024 - EPI_BYTES - 1536;   // alignment slack, static shared
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 6 ? 6 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int DYN_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float z) {
  return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
}
// erf GELU and its derivative from one shared evaluation: Phi(x) = 0.5 (1 + erf(x / sqrt 2)) with erf from Abramowitz &
// Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16 resolution), whose exp(-x^2/2) factor is also the Gaussian density
// the derivative needs:  gelu = x Phi,  gelu' = Phi + x exp(-x^2/2) / sqrt(2 pi).  ~17 FP32 ops + 2 MUFU per element,
// against erff + expf (~35) for the two separate evaluations: the fc1 / dZ epilogues are issue-bound otherwise.
__device__ __forceinline__ void gelu_fwd_grad(float x, float& g, float& d) {
  const float z = x * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, fabsf(z), 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float phi = fmaf(copysignf(0.5f, z), erf_abs, 0.5f);
  g = x * phi;
  d = fmaf(x * 0.3989422804014327f, e, phi);
}

// FAST: the epilogue is specialised for the training step's bf16 GEMMs - C, residual / aux input and aux output all go
// through TMA, no fp32 / atomic / legacy pre-activation paths - which removes two thirds of the epilogue's code (the generic
// epilogue spent 8-15 % of its issue slots waiting for instruction fetch, profiles/r1_summary.md).
template <int BLOCK_N, bool A_MN, bool B_MN, bool CG2, bool FAST>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_pre,
               const __grid_constant__ CUtensorMap tmap_x, const GemmKParams p) {
  using Cfg = SmemCfg<BLOCK_N, CG2>;
  constexpr int M_TILE = CG2 ? 2 * BLOCK_M : BLOCK_M;      // rows of C per scheduling unit
  const uint32_t rank = CG2 ? cluster_ctarank() : 0u;       // 0 = leader (issues the MMAs of the pair)
  const long long unit0 = CG2 ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long unit_stride = CG2 ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  constexpr int STAGES = Cfg::STAGES;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // 256 or 512: power of two >= 32

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_bias[BLOCK_N];      // bias of the current tile's columns (0 when absent)
  __shared__ __align__(8) uint64_t epi_bar[NUM_EPI_WARPS];   // per epilogue warp: TMA load of its residual / aux box

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  // 1024-byte aligned tile ring (required by the 128B swizzle atom = 8 rows x 128 B).
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], CG2 ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(&epi_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CG2) tmem_alloc_2cta(&tmem_base_smem, TMEM_COLS); else tmem_alloc(&tmem_base_smem, TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (CG2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  pdl_wait();   // on-chip state is set up; global memory of the preceding kernel is visible from here on
  const uint32_t tmem_base = tmem_base_smem;

  const long long units_per_batch = (long long)p.num_m_blocks * p.num_n_blocks * p.split_k;
  auto ncol0 = [&](int n_blk, int nh) { return n_blk * BLOCK_N + (nh == 2 ? BLOCK_N / 2 : 0); };

  // unit index -> (m_blk, n_blk, split, h, b); m fastest so concurrently running CTAs share a B tile.
  auto decode = [&](long long w, int& m_blk, int& n_blk, int& split, int& h, int& b, int& nh) {
    nh = 0;                                   // 0: whole BLOCK_N-wide tile, 1 / 2: its first / second column half
    if (CG2 && p.full_units >= 0 && w >= p.full_units) {
      const long long e = w - p.full_units;
      nh = 1 + (int)(e & 1);
      w = p.full_units + (e >> 1);
    }
    long long batch = w / units_per_batch;
    int r = (int)(w - batch * units_per_batch);
    m_blk = r % p.num_m_blocks;
    r /= p.num_m_blocks;
    n_blk = r % p.num_n_blocks;
    split = r / p.num_n_blocks;
    h = (int)(batch % p.batch_h);
    b = (int)(batch / p.batch_h);
  };
  auto k_range = [&](int m_blk, int n_blk, int split, int& kb0, int& kb1) {
    kb0 = split * p.kb_per_split;
    kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
    if (p.causal == 2) kb1 = min(kb1, ((m_blk + 1) * M_TILE + BLOCK_K - 1) / BLOCK_K);
    if (p.causal == 1 && n_blk * BLOCK_N > m_blk * M_TILE + M_TILE - 1) kb1 = kb0;  // fully masked tile
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long w = unit0; w < p.total_units; w += unit_stride) {
        int m_blk, n_blk, split, h, b, kb0, kb1, nh;
        decode(w, m_blk, n_blk, split, h, b, nh);
        k_range(m_blk, n_blk, split, kb0, kb1);
        // this CTA's 128 rows of A and its share of the B tile (a half-width unit uses the first B_ROWS / 2 rows of the box)
        const int m0 = m_blk * M_TILE + (int)rank * BLOCK_M, n0 = ncol0(n_blk, nh) + (int)rank * (nh ? Cfg::B_ROWS / 2 : Cfg::B_ROWS);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_gen + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int k0 = kb * BLOCK_K;
          if constexpr (!CG2) {
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if constexpr (!A_MN) {
              tma_load_4d(sa, &tmap_a, &full_bar[stage], k0, m0, h, b);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_4d(sa + j * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + 64 * j, k0, h, b);
            }
            if constexpr (!B_MN) {
              tma_load_4d(sb, &tmap_b, &full_bar[stage], k0, n0, h, b);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_4d(sb + j * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n0 + 64 * j, k0, h, b);
            }
          } else {
            // both CTAs' bytes are accounted on the LEADER's full barrier; only the leader arms it
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if constexpr (!A_MN) {
              tma_load_4d_2cta(sa, &tmap_a, lbar, k0, m0, h, b);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_4d_2cta(sa + j * (BLOCK_K * 128), &tmap_a, lbar, m0 + 64 * j, k0, h, b);
            }
            if constexpr (!B_MN) {
              tma_load_4d_2cta(sb, &tmap_b, lbar, k0, n0, h, b);
            } else {
#pragma unroll
              for (int j = 0; j < Cfg::B_ROWS / 64; ++j)
                tma_load_4d_2cta(sb + j * (BLOCK_K * 128), &tmap_b, lbar, n0 + 64 * j, k0, h, b);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      // producer tail: the release of every slot this CTA filled (a tcgen05.commit arrival, multicast to both CTAs of a
      // pair) has landed before the CTA exits and its shared memory is handed to the next CTA
      for (int i = 0; i < STAGES; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc_full = make_idesc_bf16(M_TILE, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t idesc_half = make_idesc_bf16(M_TILE, BLOCK_N / 2, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long w = unit0; w < p.total_units; w += unit_stride) {
        int m_blk, n_blk, split, h, b, kb0, kb1, nh;
        decode(w, m_blk, n_blk, split, h, b, nh);
        k_range(m_blk, n_blk, split, kb0, kb1);
        if (kb1 <= kb0) continue;  // nothing to accumulate; the epilogue skips this unit as well
        const uint32_t idesc = nh ? idesc_half : idesc_full;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            // K-major: 16 k-elements = 32 B inside the 128 B swizzle row; 8-row groups 1024 B apart.
            // MN-major: 16 k-rows of 128 B = 2048 B; 64-wide mn blocks BLOCK_K*128 B apart.
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(sa + kk * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                        : make_smem_desc_sw128(sa + kk * (UMMA_K * 2), 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(sb + kk * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                        : make_smem_desc_sw128(sb + kk * (UMMA_K * 2), 16, 1024);
            if constexpr (CG2) umma_bf16_2cta(tmem_d, adesc, bdesc, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          // smem slot is free (in both CTAs of a pair) once these MMAs retire
          if constexpr (CG2) umma_commit_2cta(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CG2) umma_commit_2cta(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;        // which half of the tile's column chunks this warp handles
    constexpr int CH_PER_WARP = BLOCK_N / 64;
    const int etid = threadIdx.x - 64;       // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    // (with FAST these are compile-time constants and the branches on them disappear)
    const int flags = FAST ? (p.flags & (PB_GEMM_GELU | PB_GEMM_AUX_DGELU | PB_GEMM_MUL_AUX)) : p.flags;
    const bool out_f32 = FAST ? false : bool(flags & PB_GEMM_OUT_F32);
    const bool do_gelu = flags & PB_GEMM_GELU;
    const bool atomic_acc = FAST ? false : bool(flags & PB_GEMM_ATOMIC_ACC);
    const bool res_f32 = FAST ? false : bool(flags & PB_GEMM_RES_F32);
    const bool tma_c = FAST ? true : (p.tma_c != 0);
    const bool tma_x = FAST ? bool(flags & PB_GEMM_AUX_DGELU) : (p.tma_x != 0);
    // TMA epilogue: a thread owns one accumulator row, so direct global accesses touch 32 different rows per warp
    // instruction (32 L1 wavefronts per request, partial sectors) and the LSU - not the tensor pipe - bounded every GEMM
    // with a residual, aux operand or second output.  Instead each warp stages its [32 rows x 32 columns] bf16 box in
    // shared memory (64-byte swizzle: conflict-free 16-byte accesses) and moves it with one bulk tensor copy.
    const int ew = warp - 2;
    uint8_t* st_buf = smem_gen + STAGES * Cfg::STAGE_BYTES + ew * Cfg::EPI_WARP_BYTES;   // slot 0 (stores)
    uint8_t* ld_buf = st_buf + 2048;                                                      // slot 1 (loads, else stores)
    const bool two_store_slots = (p.tma_pre == 0);
    uint64_t* ld_bar = &epi_bar[ew];
    uint32_t ld_phase = 0;
    int st_slot = 0;
    const int sw = (lane >> 1) & 3;          // 64-byte swizzle: 16-byte chunk g of row `lane` lives at chunk g ^ sw
    if (lane == 0) {
      if (tma_c) tma_prefetch_desc(&tmap_c);
      if (p.tma_pre) tma_prefetch_desc(&tmap_pre);
      if (tma_x) tma_prefetch_desc(&tmap_x);
    }
    for (long long w = unit0; w < p.total_units; w += unit_stride) {
      int m_blk, n_blk, split, h, b, kb0, kb1, nh;
      decode(w, m_blk, n_blk, split, h, b, nh);
      k_range(m_blk, n_blk, split, kb0, kb1);
      if (kb1 <= kb0) continue;
      const int nc0 = ncol0(n_blk, nh);                         // first column of this unit
      const int chw = nh ? CH_PER_WARP / 2 : CH_PER_WARP;       // 32-column chunks per epilogue warp
      const int row0w = m_blk * M_TILE + (int)rank * BLOCK_M + quad * 32;   // first row of this warp's 32-row block
      const int row = row0w + lane;
      const bool row_ok = row < p.M;
      const long long c_off = (long long)b * p.c_stride_b + (long long)h * p.c_stride_h + (long long)row * p.ldc;
      const long long r_off = (long long)b * p.r_stride_b + (long long)h * p.r_stride_h +
                              (long long)(p.r_row_mod > 0 ? row % p.r_row_mod : row) * p.ldr;
      const bool first_split = (split == 0);
      // stage this tile's bias slice in shared memory (read by every row, otherwise 8 dependent global loads
      // per chunk sit on the epilogue's critical path)
      asm volatile("bar.sync 1, 256;" ::: "memory");   // every warp is done with the previous tile's bias
      for (int i = etid; i < BLOCK_N; i += NUM_EPI_WARPS * 32) {
        const int col = nc0 + i;
        s_bias[i] = (p.bias != nullptr && first_split && col < p.N) ? __ldg(p.bias + col) : 0.f;
      }
      // Global operands of the epilogue (residual rows, gelu' / aux operand) are fetched one 32-column chunk ahead -
      // chunk 0 before the accumulator is even complete - so their latency never sits between the TMEM read and the
      // stores.  One prefetch stream: the aux operand when the epilogue multiplies by it, else the residual; by TMA
      // when the host could build a tensor map for it (p.tma_pre), else with per-thread vector loads.
      const bool pre_tma = p.tma_pre != 0 && row0w < p.M;
      const bool want_aux = FAST ? false : (row_ok && !out_f32 && (flags & (PB_GEMM_MUL_DGELU | PB_GEMM_MUL_AUX)) && ((p.ldaux & 7) == 0));
      const bool want_res = FAST ? false : (!want_aux && row_ok && !out_f32 && p.residual != nullptr && first_split && !res_f32 && ((p.ldr & 7) == 0));
      const __nv_bfloat16* res_row = reinterpret_cast<const __nv_bfloat16*>(p.residual) + r_off;
      const __nv_bfloat16* aux_row = reinterpret_cast<const __nv_bfloat16*>(p.aux) + (long long)row * p.ldaux;
      const __nv_bfloat16* pre_row = want_aux ? aux_row : res_row;
      auto issue_loads = [&](int chi, uint4 (&rr)[4]) {
        const int c0 = nc0 + (half * chw + chi) * 32;
        if (pre_tma) {
          if (lane == 0 && c0 < p.N) {
            mbar_expect_tx(ld_bar, 2048);
            tma_load_4d(ld_buf, &tmap_pre, ld_bar, c0, row0w, h, b);
          }
        } else if ((want_res || want_aux) && c0 + 32 <= p.N) {
          const uint4* r = reinterpret_cast<const uint4*>(pre_row + c0);
#pragma unroll
          for (int g = 0; g < 4; ++g) rr[g] = r[g];
        }
      };
      // packs 32 fp32 values to bf16 and stores them as this thread's row of the warp's box through a TMA store
      auto tma_store_row = [&](const CUtensorMap* tm, const float (&val)[32], int col0) {
        if (lane == 0) {                             // the slot about to be overwritten has been read
          if (two_store_slots) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        __syncwarp();
        uint8_t* sb = st_buf + st_slot * 2048 + lane * 64;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(val[g * 8 + 2 * t], val[g * 8 + 2 * t + 1]);
          *reinterpret_cast<uint4*>(sb + ((g ^ sw) << 4)) = o;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(tm, st_buf + st_slot * 2048, col0, row0w, h, b);
          bulk_commit();
        }
        if (two_store_slots) st_slot ^= 1;
      };
      uint4 rpre[4], npre[4];
      issue_loads(0, rpre);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll 1
      for (int chi = 0; chi < chw; ++chi) {
        const int ch = half * chw + chi;
        uint32_t v[32];
        const int col0 = nc0 + ch * 32;
        if (col0 >= p.N || row0w >= p.M) break;       // warp-uniform: nothing of this chunk (or any later one) exists
        const bool full = (col0 + 32 <= p.N);
        bool pf_res = want_res && full, pf_aux = want_aux && full;
        __syncwarp();  // tcgen05.ld is .sync.aligned
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N + ch * 32), v);
        if (pre_tma) {
          mbar_wait(ld_bar, ld_phase);
          ld_phase ^= 1;
#pragma unroll
          for (int g = 0; g < 4; ++g) rpre[g] = *reinterpret_cast<const uint4*>(ld_buf + lane * 64 + ((g ^ sw) << 4));
          __syncwarp();                               // every lane has read the slot: the next box may land in it
          if (chi + 1 < chw) issue_loads(chi + 1, npre);
          // out-of-range rows / columns of the box are zero-filled: use the staged values whenever they are requested
          pf_aux = (p.tma_pre == 2);
          pf_res = (p.tma_pre == 1) && first_split;
        } else if (chi + 1 < chw) {
          issue_loads(chi + 1, npre);
        }
        tmem_ld_wait();
        if (row_ok || tma_c || tma_x) {   // TMA stores are warp-collective: rows >= M are clipped by the tensor map
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) * p.alpha;
        {
          const float* sb = &s_bias[ch * 32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] += sb[j];
        }
        if (flags & PB_GEMM_AUX_PREACT) {
          const long long a_off = (long long)row * p.ldaux + col0;
          if (out_f32) {
            float* ax = reinterpret_cast<float*>(p.aux) + a_off;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) ax[j] = x[j];
          } else {
            __nv_bfloat16* ax = reinterpret_cast<__nv_bfloat16*>(p.aux) + a_off;
            if (tma_x) {
              tma_store_row(&tmap_x, x, col0);
            } else if (full && ((p.ldaux & 7) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 o;
                __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(x[j + 2 * t], x[j + 2 * t + 1]);
                *reinterpret_cast<uint4*>(ax + j) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) ax[j] = __float2bfloat16(x[j]);
            }
            // the activation below must see exactly what backward will re-read
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __bfloat162float(__float2bfloat16(x[j]));
          }
        }
        if (do_gelu) {
          if (flags & PB_GEMM_AUX_DGELU) {
            float dg[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) gelu_fwd_grad(x[j], x[j], dg[j]);
            __nv_bfloat16* ax = reinterpret_cast<__nv_bfloat16*>(p.aux) + (long long)row * p.ldaux + col0;
            if (tma_x) {
              tma_store_row(&tmap_x, dg, col0);
            } else if (full && ((p.ldaux & 7) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 o;
                __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(dg[j + 2 * t], dg[j + 2 * t + 1]);
                *reinterpret_cast<uint4*>(ax + j) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) ax[j] = __float2bfloat16(dg[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float dummy;
              gelu_fwd_grad(x[j], x[j], dummy);
            }
          }
        }
        if (flags & PB_GEMM_MUL_AUX) {
          const __nv_bfloat16* ax = aux_row + col0;
          if (pf_aux) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rpre[j >> 3]);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __bfloat1622float2(r2[t]);
                x[j + 2 * t] *= f.x;
                x[j + 2 * t + 1] *= f.y;
              }
            }
          } else if (!FAST && row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) x[j] *= __bfloat162float(ax[j]);
          }
        }
        if ((flags & PB_GEMM_MUL_DGELU) && (row_ok || pf_aux)) {
          const long long a_off = (long long)row * p.ldaux + col0;
          if (out_f32) {
            const float* ax = reinterpret_cast<const float*>(p.aux) + a_off;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) x[j] *= dgelu_erf(ax[j]);
          } else {
            const __nv_bfloat16* ax = reinterpret_cast<const __nv_bfloat16*>(p.aux) + a_off;
            if (pf_aux || (full && ((p.ldaux & 7) == 0))) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 rv = pf_aux ? rpre[j >> 3] : *reinterpret_cast<const uint4*>(ax + j);
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 f = __bfloat1622float2(r2[t]);
                  x[j + 2 * t] *= dgelu_erf(f.x);
                  x[j + 2 * t + 1] *= dgelu_erf(f.y);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) x[j] *= dgelu_erf(__bfloat162float(ax[j]));
            }
          }
        }
        if (p.drop.seed != nullptr) {
          const uint32_t key = pbdrop::site_key(*p.drop.seed, p.drop.op);
          const unsigned long long base = (unsigned long long)row * (unsigned long long)p.N + (unsigned long long)col0;
          if (FAST || (base & 31ull) == 0) {
            const uint32_t bits = pbdrop::keep_bits<32>(key, base, p.drop.thresh);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = ((bits >> j) & 1u) ? x[j] * p.drop.scale : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = pbdrop::keep(key, base + j, p.drop.thresh) ? x[j] * p.drop.scale : 0.f;
          }
        }
        if (p.residual != nullptr && first_split && (FAST ? pf_res : (row_ok || pf_res))) {
          if (res_f32) {
            const float* r = reinterpret_cast<const float*>(p.residual) + r_off + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) x[j] += r[j];
          } else {
            const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(p.residual) + r_off + col0;
            if (FAST || pf_res || (full && ((p.ldr & 7) == 0))) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 rv = (FAST || pf_res) ? rpre[j >> 3] : *reinterpret_cast<const uint4*>(r + j);
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 f = __bfloat1622float2(r2[t]);
                  x[j + 2 * t] += f.x;
                  x[j + 2 * t + 1] += f.y;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) x[j] += __bfloat162float(r[j]);
            }
          }
        }
        if (out_f32) {
          float* c = reinterpret_cast<float*>(p.c) + c_off + col0;
          if (atomic_acc) {
            if (full && ((p.ldc & 3) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)   // 16-byte vector reduction (red.global.add.v4.f32, sm_90+)
                atomicAdd(reinterpret_cast<float4*>(c + j), make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) atomicAdd(c + j, x[j]);
            }
          } else if (full && ((p.ldc & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(c + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) c[j] = x[j];
          }
        } else {
          __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.c) + c_off + col0;
          if (tma_c) {
            tma_store_row(&tmap_c, x, col0);
          } else if (!row_ok) {
            // (only reached when the aux output alone goes through TMA)
          } else if (full && ((p.ldc & 7) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 o;
              __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(x[j + 2 * t], x[j + 2 * t + 1]);
              *reinterpret_cast<uint4*>(c + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) c[j] = __float2bfloat16(x[j]);
          }
        }
        }  // row_ok
#pragma unroll
        for (int g = 0; g < 4; ++g) rpre[g] = npre[g];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));  // leader's barrier
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_read<0>();   // outstanding TMA stores still read this CTA's shared memory
  }

  tc_fence_before();
  if constexpr (CG2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if constexpr (CG2) tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

struct TmapKey {
  const void* base;
  uint64_t d[4];
  uint64_t s[3];
  uint32_t b0, b1;
  uint32_t swizzle, pad;
  bool operator==(const TmapKey& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    size_t hsh = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) hsh = (hsh ^ w[i]) * 1099511628211ull;
    return hsh;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mutex;

}  // namespace pb

// inner = contiguous dimension (elements); rows = second dimension; ld = row stride (elements)
int pb_make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, long long ld, int nh,
                     long long stride_h, int nb, long long stride_b, uint32_t box_inner, uint32_t box_rows) {
  return pb_make_tmap_bf16_sw(out, base, inner, rows, ld, nh, stride_h, nb, stride_b, box_inner, box_rows, 128);
}

int pb_make_tmap_bf16_sw(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, long long ld, int nh,
                        long long stride_h, int nb, long long stride_b, uint32_t box_inner, uint32_t box_rows,
                        int swizzle_bytes) {
  using namespace pb;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return pb_set_error("cuTensorMapEncodeTiled entry point not available");
  TmapKey key;
  std::memset(&key, 0, sizeof(key));
  key.base = base;
  key.d[0] = inner; key.d[1] = rows; key.d[2] = (uint64_t)nh; key.d[3] = (uint64_t)nb;
  const uint64_t full = (uint64_t)rows * (uint64_t)ld * 2ull;
  key.s[0] = (uint64_t)ld * 2ull;
  key.s[1] = nh > 1 ? (uint64_t)stride_h * 2ull : full;
  key.s[2] = nb > 1 ? (uint64_t)stride_b * 2ull : full * (uint64_t)(nh > 1 ? 1 : 1);
  key.b0 = box_inner; key.b1 = box_rows;
  key.swizzle = (uint32_t)swizzle_bytes;
  {
    std::lock_guard<std::mutex> lk(g_tmap_mutex);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *out = it->second; return 0; }
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return pb_set_error("gemm operand base not 16-byte aligned");
  for (int i = 0; i < 3; ++i)
    if (key.s[i] % 16 != 0 || key.s[i] == 0) return pb_set_error("gemm operand stride not a multiple of 16 bytes");
  cuuint64_t dims[4] = {key.d[0], key.d[1], key.d[2], key.d[3]};
  cuuint64_t strides[3] = {key.s[0], key.s[1], key.s[2]};
  cuuint32_t box[4] = {box_inner, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d) inner=%llu rows=%llu ld=%lld", (int)r,
             (unsigned long long)inner, (unsigned long long)rows, ld);
    return pb_set_error(msg);
  }
  std::lock_guard<std::mutex> lk(g_tmap_mutex);
  if (g_tmap_cache.size() > 4096) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *out);
  return 0;
}

namespace pb {
static inline int make_tmap(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, long long ld, int nh,
                            long long stride_h, int nb, long long stride_b, uint32_t box_inner, uint32_t box_rows) {
  return pb_make_tmap_bf16(out, base, inner, rows, ld, nh, stride_h, nb, stride_b, box_inner, box_rows);
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool CG2, bool FAST>
static int launch_impl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tp,
                  const CUtensorMap& tx, const GemmKParams& kp, cudaStream_t stream) {
  using Cfg = SmemCfg<BLOCK_N, CG2>;
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BLOCK_N, A_MN, B_MN, CG2, FAST>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::DYN_BYTES);
    if (e != cudaSuccess) return pb_set_cuda_error("cudaFuncSetAttribute(gemm_tc)", e);
    attr_set = true;
  }
  if constexpr (!CG2) {
    long long grid = kp.total_units < (long long)pb_num_sms() ? kp.total_units : (long long)pb_num_sms();
    PB_LAUNCH(kern, (unsigned)grid, NUM_THREADS, Cfg::DYN_BYTES, stream, ta, tb, tc, tp, tx, kp);
  } else {
    // one CTA pair (cluster of 2 = one TPC) per scheduling unit slot
    const long long pairs_max = pb_num_sms() / 2;
    const long long pairs = kp.total_units < pairs_max ? kp.total_units : pairs_max;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg::DYN_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pb_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, tp, tx, kp);
    if (e != cudaSuccess) return pb_set_cuda_error("cudaLaunchKernelEx(gemm_tc cta_group::2)", e);
  }
  return pb_check_launch("gemm_tc_kernel");
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool CG2>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tp,
                  const CUtensorMap& tx, const GemmKParams& kp, cudaStream_t stream, bool fast) {
  return fast ? launch_impl<BLOCK_N, A_MN, B_MN, CG2, true>(ta, tb, tc, tp, tx, kp, stream)
              : launch_impl<BLOCK_N, A_MN, B_MN, CG2, false>(ta, tb, tc, tp, tx, kp, stream);
}

}  // namespace pb

using namespace pb;

extern "C" int pb_gemm_bf16(const pb_gemm_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return pb_set_error("pb_gemm_bf16: empty problem");
  const int nh = d->batch_h > 0 ? d->batch_h : 1, nb = d->batch_b > 0 ? d->batch_b : 1;
  const int block_n = (d->block_n == 128 || d->block_n == 256) ? d->block_n : (d->N <= 128 ? 128 : 256);
  GemmKParams kp;
  kp.M = d->M; kp.N = d->N; kp.K = d->K;
  // cta_group::2 (CTA pairs, 256-row tiles): large non-causal problems with the 256-wide N tile
  static const int cg2_env = getenv("PIANOBART_B200_CG2") ? atoi(getenv("PIANOBART_B200_CG2")) : 1;
  bool cg2 = d->cta_group == 2 || (d->cta_group == 0 && cg2_env && d->M >= 1024);
  if (block_n != 256 || d->causal) cg2 = false;
  const int m_tile = cg2 ? 2 * BLOCK_M : BLOCK_M;
  kp.num_m_blocks = (d->M + m_tile - 1) / m_tile;
  kp.num_n_blocks = (d->N + block_n - 1) / block_n;
  kp.num_k_blocks = (d->K + BLOCK_K - 1) / BLOCK_K;
  int split = d->split_k > 1 ? d->split_k : 1;
  if (split > kp.num_k_blocks) split = kp.num_k_blocks;
  kp.kb_per_split = (kp.num_k_blocks + split - 1) / split;
  kp.split_k = (kp.num_k_blocks + kp.kb_per_split - 1) / kp.kb_per_split;
  if (kp.split_k > 1 && !((d->flags & PB_GEMM_ATOMIC_ACC) && (d->flags & PB_GEMM_OUT_F32)))
    return pb_set_error("pb_gemm_bf16: split_k > 1 needs OUT_F32|ATOMIC_ACC");
  if ((d->flags & PB_GEMM_ATOMIC_ACC) && !(d->flags & PB_GEMM_OUT_F32))
    return pb_set_error("pb_gemm_bf16: ATOMIC_ACC needs OUT_F32");
  if ((d->flags & PB_GEMM_GELU) && kp.split_k > 1) return pb_set_error("pb_gemm_bf16: GELU with split_k");
  kp.batch_h = nh; kp.batch_b = nb;
  kp.total_units = (long long)kp.num_m_blocks * kp.num_n_blocks * kp.split_k * nh * nb;
  kp.full_units = -1;
  static const int tail_env = getenv("PIANOBART_B200_TAIL_SPLIT") ? atoi(getenv("PIANOBART_B200_TAIL_SPLIT")) : 1;
  if (cg2 && tail_env && kp.split_k == 1 && nh * nb == 1 && !d->causal) {
    const long long slots = pb_num_sms() / 2, T = kp.total_units;
    const long long tail = T % slots;
    if (T > slots && tail > 0 && 2 * tail <= slots) {
      kp.full_units = (int)(T - tail);
      kp.total_units = kp.full_units + 2 * tail;
    }
  }
  kp.c = d->c; kp.ldc = d->ldc; kp.c_stride_h = d->c_stride_h; kp.c_stride_b = d->c_stride_b;
  kp.bias = d->bias;
  kp.residual = d->residual; kp.ldr = d->ldr; kp.r_stride_h = d->r_stride_h; kp.r_stride_b = d->r_stride_b;
  kp.alpha = d->alpha;
  kp.flags = d->flags;
  kp.causal = d->causal;
  kp.aux = d->aux; kp.ldaux = d->ldaux;
  kp.r_row_mod = d->r_row_mod;
  kp.drop.seed = d->drop_seed; kp.drop.op = d->drop_op; kp.drop.thresh = d->drop_thresh; kp.drop.scale = d->drop_scale;
  if (d->drop_seed != nullptr && (nh * nb != 1 || kp.split_k > 1)) return pb_set_error("pb_gemm_bf16: dropout needs no batching / split_k");
  constexpr int AUX_FLAGS = PB_GEMM_AUX_PREACT | PB_GEMM_MUL_DGELU | PB_GEMM_AUX_DGELU | PB_GEMM_MUL_AUX;
  if ((d->flags & AUX_FLAGS) && (d->aux == nullptr || nh * nb != 1 || kp.split_k > 1))
    return pb_set_error("pb_gemm_bf16: aux epilogues need aux != NULL, no batching, no split_k");
  if ((d->flags & (PB_GEMM_AUX_DGELU | PB_GEMM_MUL_AUX)) && (d->flags & PB_GEMM_OUT_F32))
    return pb_set_error("pb_gemm_bf16: AUX_DGELU / MUL_AUX are bf16-output epilogues");
  if ((d->flags & PB_GEMM_AUX_DGELU) && !(d->flags & PB_GEMM_GELU))
    return pb_set_error("pb_gemm_bf16: AUX_DGELU needs GELU");
  if (kp.causal && kp.split_k > 1) return pb_set_error("pb_gemm_bf16: causal with split_k");

  CUtensorMap ta, tb;
  int rc;
  if (!d->a_mn_major)
    rc = make_tmap(&ta, d->a, (uint64_t)d->K, (uint64_t)d->M, d->lda, nh, d->a_stride_h, nb, d->a_stride_b, 64,
                   BLOCK_M);
  else
    rc = make_tmap(&ta, d->a, (uint64_t)d->M, (uint64_t)d->K, d->lda, nh, d->a_stride_h, nb, d->a_stride_b, 64,
                   BLOCK_K);
  if (rc) return rc;
  if (!d->b_mn_major)
    rc = make_tmap(&tb, d->b, (uint64_t)d->K, (uint64_t)d->N, d->ldb, nh, d->b_stride_h, nb, d->b_stride_b, 64,
                   (uint32_t)(cg2 ? block_n / 2 : block_n));
  else
    rc = make_tmap(&tb, d->b, (uint64_t)d->N, (uint64_t)d->K, d->ldb, nh, d->b_stride_h, nb, d->b_stride_b, 64,
                   BLOCK_K);
  if (rc) return rc;

  // Epilogue tensor maps ([32 x 32] bf16 boxes, 64-byte swizzle) wherever the operand is bf16, 16-byte aligned and
  // addressed by the plain (row, column, h, b) index: C, the residual or aux input, the aux output.
  CUtensorMap tc = ta, tp = ta, tx = ta;
  static const int tma_epi_env = getenv("PIANOBART_B200_TMA_EPI") ? atoi(getenv("PIANOBART_B200_TMA_EPI")) : 1;
  auto epi_ok = [&](const void* ptr, long long ld, long long sh, long long sb) {
    if (!tma_epi_env || ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld & 7) != 0 || ld < d->N) return false;
    if (nh > 1 && ((sh & 7) != 0 || sh <= 0)) return false;
    if (nb > 1 && ((sb & 7) != 0 || sb <= 0)) return false;
    return true;
  };
  const bool out_bf16 = !(d->flags & PB_GEMM_OUT_F32);
  kp.tma_c = kp.tma_pre = kp.tma_x = 0;
  if (out_bf16 && epi_ok(d->c, d->ldc, d->c_stride_h, d->c_stride_b)) {
    if (pb_make_tmap_bf16_sw(&tc, d->c, (uint64_t)d->N, (uint64_t)d->M, d->ldc, nh, d->c_stride_h, nb, d->c_stride_b, 32, 32, 64))
      return -1;
    kp.tma_c = 1;
  }
  if (out_bf16 && (d->flags & (PB_GEMM_MUL_AUX | PB_GEMM_MUL_DGELU)) && epi_ok(d->aux, d->ldaux, 0, 0)) {
    if (pb_make_tmap_bf16_sw(&tp, d->aux, (uint64_t)d->N, (uint64_t)d->M, d->ldaux, 1, 0, 1, 0, 32, 32, 64)) return -1;
    kp.tma_pre = 2;
  } else if (out_bf16 && d->residual != nullptr && !(d->flags & PB_GEMM_RES_F32) && d->r_row_mod <= 0 && kp.split_k == 1 &&
             epi_ok(d->residual, d->ldr, d->r_stride_h, d->r_stride_b)) {
    if (pb_make_tmap_bf16_sw(&tp, d->residual, (uint64_t)d->N, (uint64_t)d->M, d->ldr, nh, d->r_stride_h, nb, d->r_stride_b,
                             32, 32, 64))
      return -1;
    kp.tma_pre = 1;
  }
  if (out_bf16 && (d->flags & (PB_GEMM_AUX_PREACT | PB_GEMM_AUX_DGELU)) && epi_ok(d->aux, d->ldaux, 0, 0)) {
    if (pb_make_tmap_bf16_sw(&tx, d->aux, (uint64_t)d->N, (uint64_t)d->M, d->ldaux, 1, 0, 1, 0, 32, 32, 64)) return -1;
    kp.tma_x = 1;
  }

  // specialised epilogue: everything the epilogue touches goes through TMA and none of the fp32 / legacy paths is needed
  static const int fast_env = getenv("PIANOBART_B200_FAST_EPI") ? atoi(getenv("PIANOBART_B200_FAST_EPI")) : 1;
  const bool fast = fast_env && kp.tma_c && kp.split_k == 1 && (d->N % 32) == 0 &&
                    !(d->flags & (PB_GEMM_OUT_F32 | PB_GEMM_ATOMIC_ACC | PB_GEMM_RES_F32 | PB_GEMM_AUX_PREACT | PB_GEMM_MUL_DGELU)) &&
                    (d->residual == nullptr || kp.tma_pre == 1) && (!(d->flags & PB_GEMM_MUL_AUX) || kp.tma_pre == 2) &&
                    (!(d->flags & PB_GEMM_AUX_DGELU) || kp.tma_x) && d->r_row_mod <= 0;

  const int variant = (d->a_mn_major ? 2 : 0) | (d->b_mn_major ? 1 : 0);
  if (cg2) {
    switch (variant) {
      case 0: return launch<256, false, false, true>(ta, tb, tc, tp, tx, kp, stream, fast);
      case 1: return launch<256, false, true, true>(ta, tb, tc, tp, tx, kp, stream, fast);
      case 2: return launch<256, true, false, true>(ta, tb, tc, tp, tx, kp, stream, fast);
      default: return launch<256, true, true, true>(ta, tb, tc, tp, tx, kp, stream, fast);
    }
  } else if (block_n == 256) {
    switch (variant) {
      case 0: return launch<256, false, false, false>(ta, tb, tc, tp, tx, kp, stream, fast);
      case 1: return launch<256, false, true, false>(ta, tb, tc, tp, tx, kp, stream, fast);
      case 2: return launch<256, true, false, false>(ta, tb, tc, tp, tx, kp, stream, fast);
      default: return launch<256, true, true, false>(ta, tb, tc, tp, tx, kp, stream, fast);
    }
  } else {
    switch (variant) {
      case 0: return launch<128, false, false, false>(ta, tb, tc, tp, tx, kp, stream, fast);
      case 1: return launch<128, false, true, false>(ta, tb, tc, tp, tx, kp, stream, fast);
      case 2: return launch<128, true, false, false>(ta, tb, tc, tp, tx, kp, stream, fast);
      default: return launch<128, true, true, false>(ta, tb, tc, tp, tx, kp, stream, fast);
    }
  }
}

// Helpers shared by the persistent decode kernels (decode_persist.cu: batch 1, decode_batch.cu: batch <= 64).
#pragma once
#include "ptx.cuh"
#include "pb_internal.h"

namespace pbdec {

typedef __nv_bfloat16 bf16;

constexpr int D = 1024, F = 2048, H = 8, HD = 128, E = 2048, V = 1280, MAXL = PB_DECODE_MAX_LAYERS;
constexpr int NSPLIT = PB_DECODE_NSPLIT, NSLOT = PB_DECODE_NSLOT;     // 18 x 57 >= 1024 keys
constexpr int NCW = 16, NCONS = NCW * 32, NTHREADS = NCONS + 32;
constexpr int SLOT_BYTES = 24576, RING = 8;
constexpr int HOPS = 3 + 10 * MAXL;                                    // tag stride per token (upper bound)
constexpr long long TIMEOUT_CYCLES = 6000000000ll;                     // ~3 s: a broken hand-off must trap, never hang

struct SampleMeta { int off[9]; float temp[8]; float top_p[8]; int pad[8]; };

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ll_store(unsigned long long* p, uint32_t payload, uint32_t tag) {
  const unsigned long long v = ((unsigned long long)tag << 32) | payload;
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void die(int* err, int code) {
  if (err) atomicExch(err, code);
  __threadfence_system();
  __trap();
}
// polls one tagged word until its tag matches
__device__ __forceinline__ uint32_t ll_wait(const unsigned long long* p, uint32_t tag, int* err) {
  unsigned long long v = ll_load(p);
  if ((uint32_t)(v >> 32) == tag) return (uint32_t)v;
  const long long t0 = clock64();
  uint32_t n = 0;
  while (true) {
    v = ll_load(p);
    if ((uint32_t)(v >> 32) == tag) return (uint32_t)v;
    if ((++n & 1023u) == 0 && clock64() - t0 > TIMEOUT_CYCLES) die(err, 2);
  }
}
__device__ __forceinline__ void mbar_wait_to(uint64_t* bar, uint32_t parity, int* err) {
  if (pb::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t n = 0;
  while (!pb::mbar_try_wait(bar, parity)) {
    if ((++n & 255u) == 0 && clock64() - t0 > TIMEOUT_CYCLES) die(err, 3);
  }
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(pb::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(pb::smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
// block reductions over the NCONS consumer threads
__device__ __forceinline__ float cons_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  cons_sync();
  if (l == 0) red[w] = v;
  cons_sync();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NCW; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float cons_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  cons_sync();
  if (l == 0) red[w] = v;
  cons_sync();
  float t = -INFINITY;
#pragma unroll
  for (int i = 0; i < NCW; ++i) t = fmaxf(t, red[i]);
  return t;
}

// PianoBartLM.sample (model.py:68-78) + sampling / nucleus (model.py:84-107) for attribute a: lg = n fp32 logits in
// shared memory (p, overwritten).  Same arithmetic as decode_sample_kernel (decode.cu).  All NCONS consumer threads call it;
// the token is returned to every thread through *tokslot.
__device__ __forceinline__ int sample_core(int tid, int n, float temp, float top_p, double u, float* p, float* sp, int* si,
                                           float* red, int* tokslot) {
  float mx = -INFINITY;
  for (int i = tid; i < n; i += NCONS) { const float x = p[i] / temp; p[i] = x; mx = fmaxf(mx, x); }
  mx = cons_max(mx, red);
  float s = 0.f;
  for (int i = tid; i < n; i += NCONS) { const float e = expf(p[i] - mx); p[i] = e; s += e; }
  s = cons_sum(s, red);
  for (int i = tid; i < n; i += NCONS) p[i] = p[i] / s;
  cons_sync();
  float part = 0.f;
  for (int i = tid; i < n; i += NCONS) part += p[i];
  const float tot1 = cons_sum(part, red) + 1e-5f;
  for (int i = tid; i < n; i += NCONS) p[i] = p[i] / tot1;
  cons_sync();
  for (int i = tid; i < n; i += NCONS) {
    const float v = p[i];
    int r = 0;
    for (int k = 0; k < n; ++k) r += (p[k] > v) || (p[k] == v && k > i);
    sp[r] = v;
    si[r] = i;
  }
  cons_sync();
  if (tid < 32) {
    const int lane = tid;
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
    if (top_p < 1.0f) {
      int first = 0x7fffffff;
#pragma unroll
      for (int i = 15; i >= 0; --i) if (k0 + i < n && excl + loc[i] > top_p) first = k0 + i;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      last = (first == 0x7fffffff) ? 1 : first + 1;
    }
    float csl = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) if (k0 + i < last) csl += sp[k0 + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) csl += __shfl_xor_sync(0xffffffffu, csl, o);
    const float cs = csl;
    double dloc[16];
    double drun = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { drun += (k0 + i < last) ? (double)(sp[k0 + i] / cs) : 0.0; dloc[i] = drun; }
    double dincl = drun;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double up = __shfl_up_sync(0xffffffffu, dincl, o); if (lane >= o) dincl += up; }
    const double dexcl = dincl - drun;
    const double dtot = __shfl_sync(0xffffffffu, dincl, 31);
    int pick = 0x7fffffff;
#pragma unroll
    for (int i = 15; i >= 0; --i) if (k0 + i < last && (dexcl + dloc[i]) / dtot > u) pick = k0 + i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pick = min(pick, __shfl_xor_sync(0xffffffffu, pick, o));
    if (pick == 0x7fffffff) pick = last - 1;
    if (lane == 0) *tokslot = si[pick];
  }
  cons_sync();
  return *tokslot;
}


}  // namespace pbdec

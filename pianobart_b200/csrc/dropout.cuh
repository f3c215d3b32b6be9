// Counter-based dropout masks (HF BART: nn.functional.dropout(p=config.dropout) after layernorm_embedding,
// after every attention out_proj and after fc2 - modeling_bart.py Bart{Encoder,Decoder}Layer.forward).
// The keep decision of element `idx` of dropout site `op` at training step `seed` is a pure function
// hash16(seed, op, idx) < threshold16, so forward and backward kernels regenerate the same mask and no mask
// tensor is ever stored.  (The reference's masks come from torch's Philox stream and cannot be reproduced
// bit for bit by any other implementation; parity tests inject this mask into the oracle instead.)
#pragma once
#include <stdint.h>

namespace pbdrop {

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t site_key(unsigned long long seed, uint32_t op) {
  return lowbias32((uint32_t)seed ^ lowbias32((uint32_t)(seed >> 32) + 0x9E3779B9U * (op + 1U)));
}
// One 32-bit hash decides a PAIR of consecutive elements (2i, 2i+1): its low / high 16 bits are compared with the top
// 16 bits of the threshold, so the keep probability is quantised to 1/65536 (p = 0.1 -> 0.100006) and the integer work
// per element halves - the hash was the largest single cost of the LayerNorm-backward and out_proj / fc2 epilogues.
__host__ __device__ __forceinline__ uint32_t pair_hash(uint32_t key, unsigned long long pair_idx) {
  return lowbias32((uint32_t)pair_idx ^ ((uint32_t)(pair_idx >> 32) * 0x85EBCA6BU) ^ key);
}
__host__ __device__ __forceinline__ bool keep(uint32_t key, unsigned long long idx, uint32_t thresh) {
  const uint32_t h = pair_hash(key, idx >> 1);
  return ((idx & 1ull) ? (h >> 16) : (h & 0xffffu)) < (thresh >> 16);
}
// keep decisions of elements idx_even and idx_even + 1 (idx_even must be even)
__host__ __device__ __forceinline__ void keep2(uint32_t key, unsigned long long idx_even, uint32_t thresh, bool& k0, bool& k1) {
  const uint32_t h = pair_hash(key, idx_even >> 1);
  k0 = (h & 0xffffu) < (thresh >> 16);
  k1 = (h >> 16) < (thresh >> 16);
}
// bit j of the result = keep decision of element idx0 + j, j < n <= 32 (idx0 even, n even): one hash per two bits
template <int NB>
__host__ __device__ __forceinline__ uint32_t keep_bits(uint32_t key, unsigned long long idx0, uint32_t thresh) {
  static_assert(NB % 2 == 0 && NB <= 32, "keep_bits");
  const uint32_t t16 = thresh >> 16;
  const uint32_t hi = (uint32_t)(idx0 >> 33) * 0x85EBCA6BU ^ key;      // pair index = idx0 >> 1; its upper word
  const uint32_t lo0 = (uint32_t)(idx0 >> 1);
  uint32_t bits = 0;
#pragma unroll
  for (int t = 0; t < NB / 2; ++t) {
    // (callers guarantee idx0 + NB does not cross a 2^33 boundary, so the upper word is constant)
    const uint32_t h = lowbias32((lo0 + t) ^ hi);
    bits |= ((h & 0xffffu) < t16 ? 1u : 0u) << (2 * t);
    bits |= ((h >> 16) < t16 ? 1u : 0u) << (2 * t + 1);
  }
  return bits;
}

struct Site {
  const unsigned long long* seed;  // device scalar, bumped once per training step
  uint32_t op;                     // dropout site id
  uint32_t thresh;                 // keep iff hash < thresh   (thresh = (1-p) * 2^32)
  float scale;                     // 1 / (1 - p)
};

}  // namespace pbdrop

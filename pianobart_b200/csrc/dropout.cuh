// Counter-based dropout masks (HF BART: nn.functional.dropout(p=config.dropout) after layernorm_embedding,
// after every attention out_proj and after fc2 - modeling_bart.py Bart{Encoder,Decoder}Layer.forward).
// The keep decision of element `idx` of dropout site `op` at training step `seed` is a pure function
// hash(seed, op, idx) < threshold, so forward and backward kernels regenerate the same mask and no mask
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
__host__ __device__ __forceinline__ bool keep(uint32_t key, unsigned long long idx, uint32_t thresh) {
  const uint32_t x = (uint32_t)idx ^ ((uint32_t)(idx >> 32) * 0x85EBCA6BU) ^ key;
  return lowbias32(x) < thresh;
}

struct Site {
  const unsigned long long* seed;  // device scalar, bumped once per training step
  uint32_t op;                     // dropout site id
  uint32_t thresh;                 // keep iff hash < thresh   (thresh = (1-p) * 2^32)
  float scale;                     // 1 / (1 - p)
};

}  // namespace pbdrop

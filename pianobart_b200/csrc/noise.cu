// BART noising as an integer kernel (reference pretrain.py:211-546 `gen_mask` + the batch prologue
// pretrain.py:128-153).  The host draws the random decisions with the reference's own RNG
// primitives (pianobart_b200/noising.py) and hands over a compact per-row plan; this kernel does
// every byte of data movement: row gather / PAD / MASK / random-token substitution, the
// "rows that changed" loss mask of SentencePermutation / TokenInfilling, the shift-right decoder
// input, the key-padding masks (computed AFTER noising, pretrain.py:151) and the int32 targets.
// One warp per row, one lane per attribute (8 active lanes x 4 rows per warp).
#include "pb_internal.h"
#include <stdint.h>

namespace {

struct NoiseConst { int pad[8]; int mask[8]; int sos[8]; };

// src codes: >= 0 source row of the same sample; -1 PAD row; -2 MASK row; <= -3 row (-3 - k) of rand_tok
// loss_mode[b]: 0 = loss flag given by the host in loss_in; 1 = loss = any(out != ori) per row;
//               2 = all-zero loss (TokenInfilling failure branch, pretrain.py:429-430)
__global__ void __launch_bounds__(256) noise_apply_kernel(const int16_t* __restrict__ ori, const int* __restrict__ src,
                                                          const int* __restrict__ rand_tok,
                                                          const uint8_t* __restrict__ loss_in,
                                                          const int* __restrict__ loss_mode, int* __restrict__ enc_ids,
                                                          int* __restrict__ dec_ids, int* __restrict__ targets,
                                                          float* __restrict__ loss_mask, uint8_t* __restrict__ enc_keep,
                                                          uint8_t* __restrict__ dec_keep, int B, int S, NoiseConst c) {
  pdl_entry();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long row = t >> 3;
  const int a = (int)(t & 7);
  if (row >= (long long)B * S) return;
  const int b = (int)(row / S), s = (int)(row % S);
  const int16_t* ob = ori + (long long)b * S * 8;
  const int o = ob[s * 8 + a];
  const int code = src[row];
  int v;
  if (code >= 0) v = ob[code * 8 + a];
  else if (code == -1) v = c.pad[a];
  else if (code == -2) v = c.mask[a];
  else v = rand_tok[(-3 - code) * 8 + a];
  enc_ids[row * 8 + a] = v;
  targets[row * 8 + a] = o;
  dec_ids[row * 8 + a] = (s == 0) ? c.sos[a] : (int)ob[(s - 1) * 8 + a];
  const int mode = loss_mode[b];
  const unsigned grp = 0xffu << ((threadIdx.x & 31) & ~7);
  const unsigned differs = __ballot_sync(0xffffffffu, v != o) & grp;
  float lm;
  if (mode == 0) lm = loss_in[row] ? 1.f : 0.f;
  else if (mode == 1) lm = differs ? 1.f : 0.f;
  else lm = 0.f;
  loss_mask[row * 8 + a] = lm;
  if (a == 0) {
    enc_keep[row] = (v != c.pad[0]) ? 1 : 0;
    const int dv = (s == 0) ? c.sos[0] : (int)ob[(s - 1) * 8];
    dec_keep[row] = (dv != c.pad[0]) ? 1 : 0;
  }
}

}  // namespace

extern "C" int pb_noise_apply(const int16_t* ori, const int* src, const int* rand_tok, const uint8_t* loss_in,
                              const int* loss_mode, int* enc_ids, int* dec_ids, int* targets, float* loss_mask,
                              uint8_t* enc_keep, uint8_t* dec_keep, int B, int S, const int* pad_host,
                              const int* mask_host, const int* sos_host, void* stream) {
  NoiseConst c;
  for (int i = 0; i < 8; ++i) { c.pad[i] = pad_host[i]; c.mask[i] = mask_host[i]; c.sos[i] = sos_host[i]; }
  const long long threads = (long long)B * S * 8;
  const int grid = (int)((threads + 255) / 256);
  PB_LAUNCH((noise_apply_kernel), grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), ori, src, rand_tok, loss_in, loss_mode,
                                                                              enc_ids, dec_ids, targets, loss_mask,
                                                                              enc_keep, dec_keep, B, S, c);
  return pb_check_launch("noise_apply");
}

// Internal helpers shared by the .cu files (error reporting, launch accounting).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/pianobart_b200.h"

int pb_set_error(const char* msg);
int pb_set_cuda_error(const char* what, cudaError_t e);
// counts the launch and converts cudaGetLastError() into the library's error convention
int pb_check_launch(const char* kernel_name);
int pb_num_sms();

// 4-D bf16 tensor map (inner, rows, h, b), 128-byte swizzle; cached by (pointer, geometry).
int pb_make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, long long ld, int nh,
                      long long stride_h, int nb, long long stride_b, uint32_t box_inner, uint32_t box_rows);
// same with an explicit swizzle span (64 or 128 bytes; box_inner * 2 must equal it)
int pb_make_tmap_bf16_sw(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, long long ld, int nh,
                         long long stride_h, int nb, long long stride_b, uint32_t box_inner, uint32_t box_rows,
                         int swizzle_bytes);

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// Kernels of one step run back to back on one stream.  Launched with programmaticStreamSerialization, a kernel's CTAs may
// be scheduled while the tail of the previous kernel is still running: they execute griddepcontrol.launch_dependents
// (let the next kernel do the same), set up their on-chip state (barriers, TMEM, descriptor prefetch) and only then
// griddepcontrol.wait - which blocks until the previous grid has completed and its memory is visible - before touching
// global memory.  Every kernel launched through PB_LAUNCH must execute pdl_wait() before its first global access.
// PIANOBART_B200_PDL=0 launches everything fully serialised.
bool pb_pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_launch_dependents(); pdl_wait(); }

template <typename... KArgs, typename... Args>
static inline cudaError_t pb_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pb_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define PB_LAUNCH(kern, grid, block, smem, stream, ...) \
  pb_launch_pdl(kern, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
#endif

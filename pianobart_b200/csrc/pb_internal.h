// Internal helpers shared by the .cu files (error reporting, launch accounting).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
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

// Internal helpers shared by the .cu files (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include "../../include/pianobart_b200.h"

int pb_set_error(const char* msg);
int pb_set_cuda_error(const char* what, cudaError_t e);
// counts the launch and converts cudaGetLastError() into the library's error convention
int pb_check_launch(const char* kernel_name);
int pb_num_sms();

// Developer microbenchmark: issue / completion cost of tcgen05.mma (kind::f16, M=128) per shape and operand source.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pianobart_b200/csrc tools/micro/mma_bench.cu -o tools/micro/_bin/mma_bench
#include "ptx.cuh"
#include <cstdio>
using namespace pb;


// mode: 0 = both operands K-major smem, 1 = B MN-major, 2 = A and B MN-major, 3 = A from TMEM (B K-major), 4 = A from TMEM, B MN-major
template <int N, int MODE, bool ALT>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_base_smem, 512);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, MODE == 2 ? 1 : 0, (MODE == 1 || MODE == 2 || MODE == 4) ? 1 : 0);
    const uint32_t aT = base, bT = base + 32768;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int kk = r & 7;
      uint64_t ad = (MODE == 2) ? make_smem_desc_sw128(aT + kk * 2048, 16384, 1024)
                                : make_smem_desc_sw128(aT + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
      uint64_t bd = (MODE == 1 || MODE == 2 || MODE == 4) ? make_smem_desc_sw128(bT + kk * 2048, 16384, 1024)
                                                          : make_smem_desc_sw128(bT + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
      const uint32_t d = tmem + ((ALT && (r & 1)) ? N : 0);
      if (MODE >= 3) umma_bf16_ts(d, tmem + 256 + kk * 8, bd, idesc, 1u);
      else umma_bf16(d, ad, bd, idesc, 1u);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int MODE, bool ALT>
void run(const char* tag, int grid) {
  long long* d; cudaMalloc(&d, 16);
  const int reps = 256;
  cudaFuncSetAttribute(bench<N, MODE, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int it = 0; it < 2; ++it) bench<N, MODE, ALT><<<grid, 128, 100 * 1024>>>(d, reps);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-34s N=%3d grid=%3d  issue %6.1f cyc/MMA   complete %6.1f cyc/MMA  (math-ideal %d)  %s\n", tag, N, grid, h[0] / (double)reps,
         h[1] / (double)reps, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<64, 0, false>("SS K-major, one accumulator", grid);
    run<64, 0, true>("SS K-major, two accumulators", grid);
    run<128, 0, false>("SS K-major, one accumulator", grid);
    run<128, 0, true>("SS K-major, two accumulators", grid);
    run<256, 0, false>("SS K-major, one accumulator", grid);
    run<64, 1, false>("SS B MN-major", grid);
    run<128, 1, false>("SS B MN-major", grid);
    run<128, 2, false>("SS A,B MN-major", grid);
    run<64, 3, false>("TS (A in TMEM) B K-major", grid);
    run<128, 3, false>("TS (A in TMEM) B K-major", grid);
    run<128, 4, false>("TS (A in TMEM) B MN-major", grid);
  }
  return 0;
}

// Error reporting and launch accounting for the C ABI (include/pianobart_b200.h).
#include "pb_internal.h"
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int pb_set_error(const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}
int pb_set_cuda_error(const char* what, cudaError_t e) {
  std::snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -2;
}
int pb_check_launch(const char* kernel_name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return pb_set_cuda_error(kernel_name, e);
  return 0;
}

extern "C" const char* pb_last_error(void) { return g_err; }
extern "C" int pb_version(void) { return 100; }
extern "C" long long pb_launch_count(void) { return g_launches.load(); }
extern "C" void pb_reset_launch_count(void) { g_launches.store(0); }

static int g_num_sms = 0;
int pb_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static std::atomic<int> g_pdl{-1};
bool pb_pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    v = (getenv("PIANOBART_B200_PDL") && atoi(getenv("PIANOBART_B200_PDL")) == 0) ? 0 : 1;
    g_pdl.store(v);
  }
  return v != 0;
}
extern "C" int pb_set_pdl(int on) {
  const int prev = pb_pdl_enabled() ? 1 : 0;
  g_pdl.store(on ? 1 : 0);
  return prev;
}

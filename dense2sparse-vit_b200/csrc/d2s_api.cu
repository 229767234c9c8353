// Error reporting, version and launch accounting for the C ABI (include/d2s.h).
#include <atomic>
#include "d2s_common.cuh"

namespace d2s {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return D2S_ERR_CUDA;
  }
  return D2S_OK;
}

}  // namespace d2s

extern "C" const char* d2s_last_error(void) { return d2s::g_err; }
extern "C" int d2s_version(void) { return 100; }
extern "C" uint64_t d2s_launch_count(void) { return d2s::g_launches.load(std::memory_order_relaxed); }

// Error state, launch counter and version for the gcl_b200 C ABI.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace gcl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail_cuda(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return GCL_ERR_CUDA;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace gcl

extern "C" int gcl_version(void) { return 4; }
extern "C" const char* gcl_last_error(void) { return gcl::g_err; }
extern "C" long long gcl_launch_count(void) { return gcl::g_launches.load(std::memory_order_relaxed); }

// Shared helpers for the gcl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gcl_b200.h"

namespace gcl {

// ---- error state (thread-local: autograd may call backward from another host thread) ----------
void set_error(const char* fmt, ...);
int fail_cuda(cudaError_t e, const char* what);

#define GCL_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::gcl::set_error(__VA_ARGS__);             \
      return GCL_ERR_BAD_ARG;                    \
    }                                            \
  } while (0)

void count_launch();

// After a launch: pick up launch-configuration errors without synchronising the stream.
#define GCL_CHECK_LAUNCH(what)                                   \
  do {                                                           \
    ::gcl::count_launch();                                       \
    cudaError_t e__ = cudaPeekAtLastError();                     \
    if (e__ != cudaSuccess) {                                    \
      cudaGetLastError();                                        \
      return ::gcl::fail_cuda(e__, what);                        \
    }                                                            \
  } while (0)

constexpr int kNumSMs = 148;  // B200

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over the `width` consecutive lanes this lane belongs to (width = 2^k <= 32)
template <int WIDTH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float prelu_f(float v, float a) { return v > 0.f ? v : a * v; }

// cp.async (LDGSTS) with zero-fill: copies `bytes` (4, 8 or 16) when pred, else writes zeros.
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gmem_src, bool pred) {
  uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  int src_size = pred ? BYTES : 0;
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_size));
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(dst), "l"(gmem_src), "n"(BYTES),
                 "r"(src_size));
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace gcl

// Single-block exclusive scan used by the graph builders (one-time work, not on the hot path).
#pragma once
#include "common.cuh"

namespace gcl {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 8;  // per thread per tile

// Single-block exclusive scan of `in[0..n)` into `out[0..n]` (out[n] = total).  in/out may alias
// only if identical pointers are NOT used (out has n+1 entries).
static __global__ void __launch_bounds__(kScanThreads) scan_exclusive_kernel(const int32_t* __restrict__ in,
                                                                      int32_t* __restrict__ out,
                                                                      int64_t n, int32_t* total_out) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t warp_off[32];
  __shared__ int32_t carry_s, tile_total;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  const int64_t tile = (int64_t)kScanThreads * kScanItems;
  for (int64_t base = 0; base < n; base += tile) {
    int32_t v[kScanItems];
    int32_t local = 0;
    const int64_t i0 = base + (int64_t)tid * kScanItems;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      v[j] = (i0 + j < n) ? in[i0 + j] : 0;
      local += v[j];
    }
    int32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int32_t w = warp_tot[lane];
      int32_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_off[lane] = wi - w;
      if (lane == 31) tile_total = wi;
    }
    __syncthreads();
    int32_t run = carry_s + warp_off[wid] + (incl - local);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      if (i0 + j < n) out[i0 + j] = run;
      run += v[j];
    }
    __syncthreads();
    if (tid == 0) carry_s += tile_total;
    __syncthreads();
  }
  if (tid == 0) {
    out[n] = carry_s;
    if (total_out) *total_out = carry_s;
  }
}

}  // namespace gcl

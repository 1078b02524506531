// K2/K3/K6: deterministic segmented SpMM over a CSR, out[b,i,:] = epi(sum_k w[k] x[b,col[k],:]).
//
// Replaces PyG's propagate (x.index_select(0, src) -> mul -> scatter_add_) for GCNConv and
// SimpleConv(mean) forward, and -- with the sender-grouped CSR -- their backward, without atomics.
// Reference call sites: /root/reference/src/models.py:414 (SimpleConv), :419 (GCNConv).
//
// Mapping: a group of L lanes (L = 4..32, power of two >= row width in 128-bit words) owns one
// (sample, row, 128-column chunk); 32/L groups share a warp.  Lanes first fetch up to L (col, w)
// pairs of the row coalesced, then walk them with group shuffles while every lane gathers one
// 128-bit word of the neighbour row per edge (4 independent gathers in flight per lane).  Samples
// are the slow grid axis so one sample's feature matrix (<= 88 MB) stays L2-resident while its rows
// are gathered deg+1 times: HBM sees each feature row once.
#include "common.cuh"

namespace gcl {
namespace {

template <int VW>
struct Vec;
template <>
struct Vec<4> {
  using T = float4;
  static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ T load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void store(float* p, T v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ void fma(T& a, float w, T v) {
    a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
  }
  static __device__ __forceinline__ T add(T a, T b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
  static __device__ __forceinline__ T prelu(T v, float s) {
    return make_float4(prelu_f(v.x, s), prelu_f(v.y, s), prelu_f(v.z, s), prelu_f(v.w, s));
  }
};
template <>
struct Vec<1> {
  using T = float;
  static __device__ __forceinline__ T zero() { return 0.f; }
  static __device__ __forceinline__ T load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void store(float* p, T v) { *p = v; }
  static __device__ __forceinline__ void fma(T& a, float w, T v) { a = fmaf(w, v, a); }
  static __device__ __forceinline__ T add(T a, T b) { return a + b; }
  static __device__ __forceinline__ T prelu(T v, float s) { return prelu_f(v, s); }
};

constexpr int kWarpsPerBlock = 8;

template <int VW, int L>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const float* __restrict__ w, const float* __restrict__ x, float* __restrict__ out,
                int64_t n_rows, int C, int nchunks, int64_t x_bstride, int64_t out_bstride,
                const float* __restrict__ bias, const float* __restrict__ prelu_slope,
                float* __restrict__ z_out) {
  using V = Vec<VW>;
  constexpr int kGroups = 32 / L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane & (L - 1), grp = lane / L;
  const unsigned mask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (grp * L));
  const int64_t item = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * kGroups + grp;
  const int64_t row = item / nchunks;
  const int chunk = (int)(item - row * nchunks);
  if (row >= n_rows) return;  // whole group leaves together (mask is per group)
  const int off = (chunk * 32 + gl) * VW;  // first column this lane owns
  const bool live = off < C;
  const float* xb = x + (int64_t)blockIdx.y * x_bstride + (live ? off : 0);

  const int32_t beg = rowptr[row], end = rowptr[row + 1];
  typename V::T acc = V::zero();
  for (int32_t base = beg; base < end; base += L) {
    const int n = min(L, end - base);
    int32_t c_reg = 0;
    float w_reg = 0.f;
    if (gl < n) {
      c_reg = __ldg(col + base + gl);
      w_reg = w ? __ldg(w + base + gl) : 1.f;
    }
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const int32_t c = __shfl_sync(mask, c_reg, j, L);
      const float wt = __shfl_sync(mask, w_reg, j, L);
      if (live) V::fma(acc, wt, V::load(xb + (int64_t)c * C));
    }
  }
  if (!live) return;
  if (bias) acc = V::add(acc, V::load(bias + off));
  const int64_t o = (int64_t)blockIdx.y * out_bstride + row * C + off;
  if (z_out) V::store(z_out + o, acc);
  if (prelu_slope) acc = V::prelu(acc, __ldg(prelu_slope));
  V::store(out + o, acc);
}

template <int VW, int L>
int launch(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, float* out, int64_t B,
           int64_t n_rows, int64_t C, int64_t xbs, int64_t obs, const float* bias, const float* slope,
           float* z_out, cudaStream_t s) {
  const int64_t units = ceil_div(C, VW);          // vector words per row
  const int nchunks = (int)ceil_div(units, 32);
  const int64_t items = n_rows * nchunks;
  const int64_t per_block = (int64_t)kWarpsPerBlock * (32 / L);
  dim3 grid((unsigned)ceil_div(items, per_block), (unsigned)B);
  spmm_kernel<VW, L><<<grid, kWarpsPerBlock * 32, 0, s>>>(rowptr, col, w, x, out, n_rows, (int)C, nchunks, xbs,
                                                         obs, bias, slope, z_out);
  GCL_CHECK_LAUNCH("gcl_spmm_f32");
  return GCL_OK;
}

template <int VW>
int dispatch_l(int64_t units, const int32_t* rowptr, const int32_t* col, const float* w, const float* x,
               float* out, int64_t B, int64_t n_rows, int64_t C, int64_t xbs, int64_t obs, const float* bias,
               const float* slope, float* z_out, cudaStream_t s) {
  if (units <= 4) return launch<VW, 4>(rowptr, col, w, x, out, B, n_rows, C, xbs, obs, bias, slope, z_out, s);
  if (units <= 8) return launch<VW, 8>(rowptr, col, w, x, out, B, n_rows, C, xbs, obs, bias, slope, z_out, s);
  if (units <= 16) return launch<VW, 16>(rowptr, col, w, x, out, B, n_rows, C, xbs, obs, bias, slope, z_out, s);
  return launch<VW, 32>(rowptr, col, w, x, out, B, n_rows, C, xbs, obs, bias, slope, z_out, s);
}

}  // namespace
}  // namespace gcl

using namespace gcl;

extern "C" int gcl_spmm_f32(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, float* out,
                            int64_t batch, int64_t n_rows_out, int64_t channels, int64_t x_bstride,
                            int64_t out_bstride, const float* bias, const float* prelu_slope, float* z_out,
                            void* stream) {
  GCL_CHECK_ARG(rowptr && col && x && out, "gcl_spmm_f32: null pointer argument");
  GCL_CHECK_ARG(x != out, "gcl_spmm_f32: x and out must not alias");
  GCL_CHECK_ARG(batch >= 0 && n_rows_out >= 0 && channels > 0 && channels < (1 << 20),
                "gcl_spmm_f32: bad sizes (batch %lld rows %lld channels %lld)", (long long)batch,
                (long long)n_rows_out, (long long)channels);
  GCL_CHECK_ARG(batch <= 65535, "gcl_spmm_f32: batch %lld exceeds 65535", (long long)batch);
  if (batch == 0 || n_rows_out == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec = (channels % 4 == 0) && (x_bstride % 4 == 0) && (out_bstride % 4 == 0) && al16(x) && al16(out) &&
                   (!bias || al16(bias)) && (!z_out || al16(z_out));
  if (vec)
    return dispatch_l<4>(channels / 4, rowptr, col, w, x, out, batch, n_rows_out, channels, x_bstride,
                         out_bstride, bias, prelu_slope, z_out, s);
  return dispatch_l<1>(channels, rowptr, col, w, x, out, batch, n_rows_out, channels, x_bstride, out_bstride,
                       bias, prelu_slope, z_out, s);
}

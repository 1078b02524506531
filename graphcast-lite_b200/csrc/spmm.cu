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

// SB samples of the same row per lane group: the row's (col, w) pairs are fetched once and every edge
// issues SB independent 128-bit gathers (one per sample), which is what hides the L2/HBM latency.
// PW: per-sample weights w[b][k] (w_bstride floats apart) times w_scale -- the attention coefficients of GATConv;
// otherwise one weight per CSR entry shared by all samples (GCN / mean).
template <int VW, int L, int SB, bool PW>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const float* __restrict__ w, const float* __restrict__ x, float* __restrict__ out,
                int64_t n_rows, int64_t n_in, int C, int nchunks, int64_t x_bstride, int64_t out_bstride, int B,
                const float* __restrict__ bias, const float* __restrict__ prelu_slope,
                float* __restrict__ z_out, int64_t w_bstride, float w_scale) {
  using V = Vec<VW>;
  constexpr int kGroups = 32 / L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane & (L - 1), grp = lane / L;
  const unsigned mask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (grp * L));
  const int64_t item = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * kGroups + grp;
  const int64_t row = item / nchunks;
  const int chunk = (int)(item - row * nchunks);
  if (row >= n_rows) return;  // whole group leaves together (mask is per group)
  const int b0 = blockIdx.y * SB;
  const int nb = min(SB, B - b0);  // block-uniform
  const int off = (chunk * 32 + gl) * VW;  // first column this lane owns
  const bool live = off < C;
  const float* xb = x + (int64_t)b0 * x_bstride + (live ? off : 0);

  const int32_t beg = rowptr[row], end = rowptr[row + 1];
  typename V::T acc[SB];
#pragma unroll
  for (int s = 0; s < SB; ++s) acc[s] = V::zero();
  for (int32_t base = beg; base < end; base += L) {
    const int n = min(L, end - base);
    int32_t c_reg = 0;
    float w_reg[PW ? SB : 1];
#pragma unroll
    for (int s = 0; s < (PW ? SB : 1); ++s) w_reg[s] = 0.f;
    if (gl < n) {
      c_reg = __ldg(col + base + gl);
      if (PW) {
#pragma unroll
        for (int s = 0; s < SB; ++s)
          if (s < nb) w_reg[s] = w[(int64_t)(b0 + s) * w_bstride + base + gl] * w_scale;
      } else {
        w_reg[0] = w ? __ldg(w + base + gl) : 1.f;
      }
      if (c_reg >= n_in) {   // entry points past the rows x holds (output restricted to a row prefix): a zero row
        c_reg = 0;
#pragma unroll
        for (int s = 0; s < (PW ? SB : 1); ++s) w_reg[s] = 0.f;
      }
    }
#pragma unroll 2
    for (int j = 0; j < n; ++j) {
      const int32_t c = __shfl_sync(mask, c_reg, j, L);
      float wt[PW ? SB : 1];
#pragma unroll
      for (int s = 0; s < (PW ? SB : 1); ++s) wt[s] = __shfl_sync(mask, w_reg[s], j, L);
      if (live) {
        typename V::T v[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s)
          if (s < nb) v[s] = V::load(xb + (int64_t)s * x_bstride + (int64_t)c * C);
#pragma unroll
        for (int s = 0; s < SB; ++s)
          if (s < nb) V::fma(acc[s], wt[PW ? s : 0], v[s]);
      }
    }
  }
  if (!live) return;
  typename V::T bv = V::zero();
  if (bias) bv = V::load(bias + off);
  const float slope = prelu_slope ? __ldg(prelu_slope) : 0.f;
#pragma unroll
  for (int s = 0; s < SB; ++s) {
    if (s >= nb) break;
    typename V::T a = V::add(acc[s], bv);
    const int64_t o = (int64_t)(b0 + s) * out_bstride + row * C + off;
    if (z_out) V::store(z_out + o, a);
    if (prelu_slope) a = V::prelu(a, slope);
    V::store(out + o, a);
  }
}

// samples per lane group: as many as keep SB feature matrices resident in L2 (126 MB) together
// Graphs with < 4 entries per row (grid<->mesh: mostly the self loop) have almost no reuse to protect,
// so they always take 8 samples per group.
inline int pick_sb(int64_t B, int64_t n_rows, int64_t C, int64_t nnz) {
  const int64_t per_sample = n_rows * C * 4;
  const bool low_reuse = nnz > 0 && nnz < 4 * n_rows;
  int sb = 8;
  while (sb > 1 && (sb > B || (!low_reuse && sb * per_sample > (48ll << 20)))) sb >>= 1;
  return sb;
}

template <int VW, int L, int SB>
int launch(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, float* out, int64_t B,
           int64_t n_rows, int64_t n_in, int64_t C, int64_t xbs, int64_t obs, const float* bias, const float* slope,
           float* z_out, int64_t wbs, float wsc, cudaStream_t s) {
  const int64_t units = ceil_div(C, VW);          // vector words per row
  const int nchunks = (int)ceil_div(units, 32);
  const int64_t items = n_rows * nchunks;
  const int64_t per_block = (int64_t)kWarpsPerBlock * (32 / L);
  dim3 grid((unsigned)ceil_div(items, per_block), (unsigned)ceil_div(B, SB));
  if (wbs > 0)
    spmm_kernel<VW, L, SB, true><<<grid, kWarpsPerBlock * 32, 0, s>>>(rowptr, col, w, x, out, n_rows, n_in, (int)C,
                                                                     nchunks, xbs, obs, (int)B, bias, slope, z_out, wbs,
                                                                     wsc);
  else
    spmm_kernel<VW, L, SB, false><<<grid, kWarpsPerBlock * 32, 0, s>>>(rowptr, col, w, x, out, n_rows, n_in, (int)C,
                                                                      nchunks, xbs, obs, (int)B, bias, slope, z_out, 0,
                                                                      1.f);
  GCL_CHECK_LAUNCH("gcl_spmm_f32");
  return GCL_OK;
}

template <int VW, int L>
int dispatch_sb(int sb, const int32_t* rowptr, const int32_t* col, const float* w, const float* x, float* out,
                int64_t B, int64_t n_rows, int64_t n_in, int64_t C, int64_t xbs, int64_t obs, const float* bias,
                const float* slope, float* z_out, int64_t wbs, float wsc, cudaStream_t s) {
  if (sb >= 8) return launch<VW, L, 8>(rowptr, col, w, x, out, B, n_rows, n_in, C, xbs, obs, bias, slope, z_out, wbs, wsc, s);
  if (sb >= 4) return launch<VW, L, 4>(rowptr, col, w, x, out, B, n_rows, n_in, C, xbs, obs, bias, slope, z_out, wbs, wsc, s);
  if (sb >= 2) return launch<VW, L, 2>(rowptr, col, w, x, out, B, n_rows, n_in, C, xbs, obs, bias, slope, z_out, wbs, wsc, s);
  return launch<VW, L, 1>(rowptr, col, w, x, out, B, n_rows, n_in, C, xbs, obs, bias, slope, z_out, wbs, wsc, s);
}

template <int VW>
int dispatch_l(int64_t units, const int32_t* rowptr, const int32_t* col, const float* w, const float* x,
               float* out, int64_t B, int64_t n_rows, int64_t n_in, int64_t C, int64_t xbs, int64_t obs, const float* bias,
               const float* slope, float* z_out, int64_t nnz, int64_t wbs, float wsc, cudaStream_t s) {
  int sb = pick_sb(B, n_rows, C, nnz);
  static const int pw_cap = getenv("GCL_PW_SB") ? atoi(getenv("GCL_PW_SB")) : 4;
  if (wbs > 0 && sb > pw_cap) sb = pw_cap;   // per-sample weights cost SB registers + shuffles per neighbour
  if (units <= 4) return dispatch_sb<VW, 4>(sb, rowptr, col, w, x, out, B, n_rows, n_in, C, xbs, obs, bias, slope, z_out, wbs, wsc, s);
  if (units <= 8) return dispatch_sb<VW, 8>(sb, rowptr, col, w, x, out, B, n_rows, n_in, C, xbs, obs, bias, slope, z_out, wbs, wsc, s);
  if (units <= 16) return dispatch_sb<VW, 16>(sb, rowptr, col, w, x, out, B, n_rows, n_in, C, xbs, obs, bias, slope, z_out, wbs, wsc, s);
  return dispatch_sb<VW, 32>(sb, rowptr, col, w, x, out, B, n_rows, n_in, C, xbs, obs, bias, slope, z_out, wbs, wsc, s);
}

}  // namespace

// shared entry for gcl_spmm_f32 and the single-head GATConv aggregation (per-sample weights, w_bstride > 0)
int spmm_run(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, float* out, int64_t batch,
             int64_t n_rows_out, int64_t n_rows_in, int64_t channels, int64_t x_bstride, int64_t out_bstride,
             const float* bias, const float* prelu_slope, float* z_out, int64_t nnz, int64_t w_bstride, float w_scale,
             cudaStream_t s) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec = (channels % 4 == 0) && (x_bstride % 4 == 0) && (out_bstride % 4 == 0) && al16(x) && al16(out) &&
                   (!bias || al16(bias)) && (!z_out || al16(z_out));
  if (vec)
    return dispatch_l<4>(channels / 4, rowptr, col, w, x, out, batch, n_rows_out, n_rows_in, channels, x_bstride,
                         out_bstride, bias, prelu_slope, z_out, nnz, w_bstride, w_scale, s);
  return dispatch_l<1>(channels, rowptr, col, w, x, out, batch, n_rows_out, n_rows_in, channels, x_bstride, out_bstride,
                       bias, prelu_slope, z_out, nnz, w_bstride, w_scale, s);
}
}  // namespace gcl

using namespace gcl;

extern "C" int gcl_spmm_f32(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, float* out,
                            int64_t batch, int64_t n_rows_out, int64_t n_rows_in, int64_t channels, int64_t x_bstride,
                            int64_t out_bstride, const float* bias, const float* prelu_slope, float* z_out,
                            int64_t nnz, void* stream) {
  GCL_CHECK_ARG(rowptr && col && x && out, "gcl_spmm_f32: null pointer argument");
  GCL_CHECK_ARG(x != out, "gcl_spmm_f32: x and out must not alias");
  GCL_CHECK_ARG(batch >= 0 && n_rows_out >= 0 && channels > 0 && channels < (1 << 20),
                "gcl_spmm_f32: bad sizes (batch %lld rows %lld channels %lld)", (long long)batch,
                (long long)n_rows_out, (long long)channels);
  GCL_CHECK_ARG(n_rows_in > 0, "gcl_spmm_f32: n_rows_in must be positive");
  GCL_CHECK_ARG(batch <= 65535, "gcl_spmm_f32: batch %lld exceeds 65535", (long long)batch);
  if (batch == 0 || n_rows_out == 0) return GCL_OK;
  return spmm_run(rowptr, col, w, x, out, batch, n_rows_out, n_rows_in, channels, x_bstride, out_bstride, bias,
                  prelu_slope, z_out, nnz, 0, 1.f, static_cast<cudaStream_t>(stream));
}

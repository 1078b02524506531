// PReLU and LayerNorm(mode="node"), forward and backward, fp32, deterministic reductions.
//   torch.nn.PReLU (one slope)             /root/reference/src/models.py:78,90,159,316
//   torch_geometric.nn.LayerNorm(mode=node) /root/reference/src/models.py:103,370  == F.layer_norm over C
#include "common.cuh"

namespace gcl {
namespace {

constexpr int kEwThreads = 256;
constexpr int kEwMaxBlocks = 8 * kNumSMs;

__global__ void prelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ slope,
                                 float* __restrict__ y, int64_t n) {
  const float a = __ldg(slope);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) y[i] = prelu_f(x[i], a);
}

// dx = dy * (x > 0 ? 1 : a); part[blk] = sum over the block's elements of dy * x * [x <= 0]
// Each block owns a CONTIGUOUS element range so the summation order is fixed.
__global__ void prelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                 const float* __restrict__ slope, float* __restrict__ dx,
                                 float* __restrict__ part, int64_t n, int64_t per_block) {
  __shared__ float sm[kEwThreads / 32];
  const float a = __ldg(slope);
  const int64_t beg = (int64_t)blockIdx.x * per_block, end = min(n, beg + per_block);
  float s = 0.f;
  for (int64_t i = beg + threadIdx.x; i < end; i += kEwThreads) {
    const float xv = x[i], g = dy[i];
    const bool pos = xv > 0.f;
    dx[i] = pos ? g : a * g;
    s += pos ? 0.f : g * xv;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kEwThreads / 32; ++k) t += sm[k];
    part[blockIdx.x] = t;
  }
}

// dx = dy * (x > 0 ? 1 : a) in one pass that ALSO produces the column sums of dx (the bias gradient of the
// layer whose PReLU this is) and the slope gradient: part[blk][0..C) = column partials, part[blk][C] = slope.
// Block (32, 8): lanes run along columns, 8 row lanes; each block owns a contiguous row range.
__global__ void __launch_bounds__(256)
    prelu_bwd_colsum_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                            const float* __restrict__ slope, float* __restrict__ dx, float* __restrict__ part,
                            int64_t rows, int C, int64_t rows_per_block) {
  __shared__ float sm[8][33];
  __shared__ float ss[8];
  const float a = __ldg(slope);
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float sl = 0.f;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + tx;
    float cs = 0.f;
    if (c < C) {
      for (int64_t r = r0 + ty; r < r1; r += 8) {
        const float xv = x[r * C + c], g = dy[r * C + c];
        const bool pos = xv > 0.f;
        const float d = pos ? g : a * g;
        dx[r * C + c] = d;
        cs += d;
        sl += pos ? 0.f : g * xv;
      }
    }
    sm[ty][tx] = cs;
    __syncthreads();
    if (ty == 0 && c < C) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[k][tx];
      part[(int64_t)blockIdx.x * (C + 1) + c] = t;
    }
    __syncthreads();
  }
  sl = warp_sum(sl);
  if (tx == 0) ss[ty] = sl;
  __syncthreads();
  if (tx == 0 && ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += ss[k];
    part[(int64_t)blockIdx.x * (C + 1) + C] = t;
  }
}

// 128-bit version (C % 4 == 0, C <= 1024, 16-byte aligned rows): a thread owns one float4 column group and one of
// 256 / (C / 4) row lanes, and keeps 4 rows (8 independent 16-byte loads) in flight -- the scalar kernel above
// has a single dependent load per thread and ran at ~0.4 of the HBM roofline.
__global__ void __launch_bounds__(256)
    prelu_bwd_colsum_v4_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                               const float* __restrict__ slope, float4* __restrict__ dx, float* __restrict__ part,
                               int64_t rows, int C, int64_t rows_per_block) {
  __shared__ float4 sm[256];
  __shared__ float ss[8];
  const float a = __ldg(slope);
  const int tid = threadIdx.x, ncol = C >> 2, lanes = 256 / ncol;
  const int ci = tid % ncol, rl = tid / ncol;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  float sl = 0.f;
  if (rl < lanes) {
    for (int64_t r = r0 + rl; r < r1; r += 4 * lanes) {
      float4 xv[4], g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t rr = r + (int64_t)u * lanes;
        if (rr < r1) {
          xv[u] = __ldcs(x + rr * ncol + ci);
          g[u] = __ldcs(dy + rr * ncol + ci);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t rr = r + (int64_t)u * lanes;
        if (rr < r1) {
          float4 d;
          d.x = xv[u].x > 0.f ? g[u].x : a * g[u].x;
          d.y = xv[u].y > 0.f ? g[u].y : a * g[u].y;
          d.z = xv[u].z > 0.f ? g[u].z : a * g[u].z;
          d.w = xv[u].w > 0.f ? g[u].w : a * g[u].w;
          dx[rr * ncol + ci] = d;
          cs.x += d.x; cs.y += d.y; cs.z += d.z; cs.w += d.w;
          sl += (xv[u].x > 0.f ? 0.f : g[u].x * xv[u].x) + (xv[u].y > 0.f ? 0.f : g[u].y * xv[u].y) +
                (xv[u].z > 0.f ? 0.f : g[u].z * xv[u].z) + (xv[u].w > 0.f ? 0.f : g[u].w * xv[u].w);
        }
      }
    }
  }
  sm[tid] = cs;
  sl = warp_sum(sl);
  if ((tid & 31) == 0) ss[tid >> 5] = sl;
  __syncthreads();
  if (tid < ncol) {
    float4 t = sm[tid];
    for (int l = 1; l < lanes; ++l) {
      const float4 o = sm[l * ncol + tid];
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    float* dst = part + (int64_t)blockIdx.x * (C + 1) + 4 * tid;
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
  }
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += ss[k];
    part[(int64_t)blockIdx.x * (C + 1) + C] = t;
  }
}

__global__ void __launch_bounds__(1024) sum_partials_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
  // one block, fixed order: thread-strided partial sums, a shuffle tree per warp, then the 32 warp sums in order
  __shared__ float ws[32];
  float s = 0.f;
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += 1024) s += part[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += ws[k];
    *out = t;
  }
}

// ---- LayerNorm over the last dim, one warp per row, row cached in registers (C <= 32 * NPER) -------
template <int NPER>
__global__ void __launch_bounds__(256)
    layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean_out,
                         float* __restrict__ rstd_out, int64_t rows, int C, float eps) {
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * C;
  float v[NPER];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NPER; ++j) {
    const int c = lane + 32 * j;
    v[j] = c < C ? xr[c] : 0.f;
    s += v[j];
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NPER; ++j) {
    const int c = lane + 32 * j;
    const float d = c < C ? v[j] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  float* yr = y + row * C;
#pragma unroll
  for (int j = 0; j < NPER; ++j) {
    const int c = lane + 32 * j;
    if (c < C) {
      float o = (v[j] - mean) * rstd;
      if (gamma) o = o * __ldg(gamma + c) + __ldg(beta + c);
      yr[c] = o;
    }
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
// Each block owns a contiguous row range; its column partials of dgamma / dbeta go to
// part[blk][0][c] / part[blk][1][c] (fixed-order reduction afterwards).
template <int NPER>
__global__ void __launch_bounds__(256)
    layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                         const float* __restrict__ gamma, const float* __restrict__ mean,
                         const float* __restrict__ rstd, float* __restrict__ dx, float* __restrict__ part,
                         int64_t rows, int C, int64_t rows_per_block) {
  extern __shared__ float sm[];  // [8 warps][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float dg[NPER], db[NPER], gm[NPER];
#pragma unroll
  for (int j = 0; j < NPER; ++j) {
    dg[j] = db[j] = 0.f;
    const int c = lane + 32 * j;
    gm[j] = (gamma && c < C) ? __ldg(gamma + c) : 1.f;
  }
  for (int64_t row = r0 + warp; row < r1; row += 8) {
    const float mu = mean[row], rs = rstd[row];
    float xh[NPER], g[NPER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
      const int c = lane + 32 * j;
      const bool in = c < C;
      const float d = in ? dy[row * C + c] : 0.f;
      xh[j] = in ? (x[row * C + c] - mu) * rs : 0.f;
      g[j] = d * gm[j];
      dg[j] += d * xh[j];
      db[j] += d;
      s1 += g[j];
      s2 += g[j] * xh[j];
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
      const int c = lane + 32 * j;
      if (c < C) dx[row * C + c] = rs * (g[j] - s1 - xh[j] * s2);
    }
  }
  if (part == nullptr) return;
#pragma unroll
  for (int j = 0; j < NPER; ++j) {
    const int c = lane + 32 * j;
    if (c < C) {
      sm[(warp * 2 + 0) * C + c] = dg[j];
      sm[(warp * 2 + 1) * C + c] = db[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w * 2 * C + i];
    part[(int64_t)blockIdx.x * 2 * C + i] = t;
  }
}

// ---- 128-bit LayerNorm (C % 4 == 0, C <= 128, 16-byte aligned rows) ---------------------------------------------
// L = 8 / 16 / 32 lanes per row, a lane holds one float4 of it; a block owns a contiguous row range and each lane
// group keeps two rows in flight.  (The warp-per-row scalar kernels below ran at ~0.4 of the HBM roofline: one
// 4-byte load per lane and row, and a warp retired after a single row.)
template <int L>
__device__ __forceinline__ float lsum(float v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, L);
  return v;
}
__device__ __forceinline__ float hsum4(float4 v) { return (v.x + v.y) + (v.z + v.w); }

template <int L>
__global__ void __launch_bounds__(256)
    layernorm_fwd_v4_kernel(const float4* __restrict__ x, const float* __restrict__ gamma,
                            const float* __restrict__ beta, float4* __restrict__ y, float* __restrict__ mean_out,
                            float* __restrict__ rstd_out, int64_t rows, int C, float eps, int64_t rows_per_block) {
  constexpr int kGroups = 256 / L;
  const int tid = threadIdx.x, gl = tid & (L - 1), grp = tid / L, ncol = C >> 2;
  const bool live = gl < ncol;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gamma && live) {
    g4 = __ldg(reinterpret_cast<const float4*>(gamma) + gl);
    b4 = __ldg(reinterpret_cast<const float4*>(beta) + gl);
  }
  const float inv_c = 1.f / (float)C;
  for (int64_t base = r0; base < r1; base += 2 * kGroups) {   // block-uniform trip count
    float4 v[2];
    int64_t row[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      row[u] = base + u * kGroups + grp;
      v[u] = (live && row[u] < r1) ? __ldcs(x + row[u] * ncol + gl) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float mean = lsum<L>(hsum4(v[u])) * inv_c;
      const float4 d = live ? make_float4(v[u].x - mean, v[u].y - mean, v[u].z - mean, v[u].w - mean)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      const float rstd = rsqrtf(lsum<L>((d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w)) * inv_c + eps);
      if (row[u] < r1) {
        if (live)
          y[row[u] * ncol + gl] = make_float4(d.x * rstd * g4.x + b4.x, d.y * rstd * g4.y + b4.y,
                                              d.z * rstd * g4.z + b4.z, d.w * rstd * g4.w + b4.w);
        if (gl == 0) {
          if (mean_out) mean_out[row[u]] = mean;
          if (rstd_out) rstd_out[row[u]] = rstd;
        }
      }
    }
  }
}

template <int L>
__global__ void __launch_bounds__(256)
    layernorm_bwd_v4_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                            const float* __restrict__ gamma, const float* __restrict__ mean,
                            const float* __restrict__ rstd, float4* __restrict__ dx, float* __restrict__ part,
                            int64_t rows, int C, int64_t rows_per_block) {
  constexpr int kGroups = 256 / L;
  __shared__ float4 sm[2][256];
  const int tid = threadIdx.x, gl = tid & (L - 1), grp = tid / L, ncol = C >> 2;
  const bool live = gl < ncol;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  const float4 gm = (gamma && live) ? __ldg(reinterpret_cast<const float4*>(gamma) + gl)
                                    : make_float4(1.f, 1.f, 1.f, 1.f);
  const float inv_c = 1.f / (float)C;
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
  for (int64_t base = r0; base < r1; base += 2 * kGroups) {
    float4 d[2], xv[2];
    float mu[2], rs[2];
    int64_t row[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      row[u] = base + u * kGroups + grp;
      const bool ok = live && row[u] < r1;
      d[u] = ok ? __ldcs(dy + row[u] * ncol + gl) : make_float4(0.f, 0.f, 0.f, 0.f);
      xv[u] = ok ? __ldcs(x + row[u] * ncol + gl) : make_float4(0.f, 0.f, 0.f, 0.f);
      mu[u] = row[u] < r1 ? __ldg(mean + row[u]) : 0.f;
      rs[u] = row[u] < r1 ? __ldg(rstd + row[u]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool ok = live && row[u] < r1;
      const float4 xh = ok ? make_float4((xv[u].x - mu[u]) * rs[u], (xv[u].y - mu[u]) * rs[u],
                                         (xv[u].z - mu[u]) * rs[u], (xv[u].w - mu[u]) * rs[u])
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 g = make_float4(d[u].x * gm.x, d[u].y * gm.y, d[u].z * gm.z, d[u].w * gm.w);
      dg.x += d[u].x * xh.x; dg.y += d[u].y * xh.y; dg.z += d[u].z * xh.z; dg.w += d[u].w * xh.w;
      db.x += d[u].x; db.y += d[u].y; db.z += d[u].z; db.w += d[u].w;
      const float s1 = lsum<L>(hsum4(g)) * inv_c;
      const float s2 = lsum<L>((g.x * xh.x + g.y * xh.y) + (g.z * xh.z + g.w * xh.w)) * inv_c;
      if (ok)
        dx[row[u] * ncol + gl] = make_float4(rs[u] * (g.x - s1 - xh.x * s2), rs[u] * (g.y - s1 - xh.y * s2),
                                             rs[u] * (g.z - s1 - xh.z * s2), rs[u] * (g.w - s1 - xh.w * s2));
    }
  }
  if (part == nullptr) return;
  sm[0][tid] = dg;
  sm[1][tid] = db;
  __syncthreads();
  if (tid < 2 * ncol) {
    const int which = tid / ncol, cc = tid % ncol;
    float4 t = sm[which][cc];
    for (int g = 1; g < kGroups; ++g) {
      const float4 o = sm[which][g * L + cc];
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    float* dst = part + ((int64_t)blockIdx.x * 2 + which) * C + 4 * cc;
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
  }
}

// out[i] = sum_k part[k * n + i] in a fixed order (32 interleaved partial sums, then those 32 ascending); block
// (32, 32) per 32 outputs.  Columns < C go to out0, the rest to out1.
__global__ void reduce_cols_kernel(const float* __restrict__ part, int nblk, int n, float* __restrict__ out0,
                                   float* __restrict__ out1, int C) {
  __shared__ float sm[32][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (i < n)
#pragma unroll 4
    for (int k = threadIdx.y; k < nblk; k += 32) s += part[(int64_t)k * n + i];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    float t = sm[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < 32; ++y) t += sm[y][threadIdx.x];
    if (i < C) out0[i] = t;
    else out1[i - C] = t;
  }
}

struct LnPlan {
  int nblk;
  int64_t rows_per_block;
};
LnPlan ln_plan(int64_t rows) {
  int64_t nblk = 4 * kNumSMs;
  int64_t rpb = ceil_div(rows, nblk);
  if (rpb < 8) rpb = 8;
  nblk = ceil_div(rows, rpb);
  if (nblk < 1) nblk = 1;
  return {(int)nblk, rpb};
}

int ew_blocks(int64_t n) {
  int64_t b = ceil_div(n, kEwThreads);
  return (int)(b < 1 ? 1 : (b > kEwMaxBlocks ? kEwMaxBlocks : b));
}

}  // namespace
}  // namespace gcl

using namespace gcl;

extern "C" int gcl_prelu_fwd_f32(const float* x, const float* slope, float* y, int64_t n, void* stream) {
  GCL_CHECK_ARG(x && slope && y && n >= 0, "gcl_prelu_fwd_f32: bad argument");
  if (n == 0) return GCL_OK;
  prelu_fwd_kernel<<<ew_blocks(n), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, slope, y, n);
  GCL_CHECK_LAUNCH("gcl_prelu_fwd_f32");
  return GCL_OK;
}

extern "C" size_t gcl_prelu_bwd_workspace_bytes(int64_t n) {
  (void)n;
  return (size_t)kEwMaxBlocks * sizeof(float) + 256;
}

extern "C" int gcl_prelu_bwd_f32(const float* dy, const float* x, const float* slope, float* dx, float* dslope,
                                 int64_t n, void* workspace, size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(dy && x && slope && dx && dslope && workspace && n >= 0, "gcl_prelu_bwd_f32: bad argument");
  if (workspace_bytes < gcl_prelu_bwd_workspace_bytes(n)) {
    set_error("gcl_prelu_bwd_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    cudaMemsetAsync(dslope, 0, sizeof(float), s);
    return GCL_OK;
  }
  const int nblk = ew_blocks(n);
  const int64_t per_block = ceil_div(n, nblk);
  float* part = static_cast<float*>(workspace);
  prelu_bwd_kernel<<<nblk, kEwThreads, 0, s>>>(dy, x, slope, dx, part, n, per_block);
  GCL_CHECK_LAUNCH("gcl_prelu_bwd_f32");
  sum_partials_kernel<<<1, 1024, 0, s>>>(part, nblk, dslope);
  GCL_CHECK_LAUNCH("gcl_prelu_bwd_f32(reduce)");
  return GCL_OK;
}

extern "C" size_t gcl_prelu_bwd_colsum_workspace_bytes(int64_t rows, int64_t c) {
  if (rows < 0 || c <= 0) return 0;
  return (size_t)ln_plan(rows).nblk * (size_t)(c + 1) * sizeof(float) + 256;
}

extern "C" int gcl_prelu_bwd_colsum_f32(const float* dy, const float* x, const float* slope, float* dx, float* dslope,
                                        float* dbias, int64_t rows, int64_t c, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(dy && x && slope && dx && dslope && dbias && workspace && rows >= 0 && c > 0 && c < (1 << 20),
                "gcl_prelu_bwd_colsum_f32: bad argument");
  if (workspace_bytes < gcl_prelu_bwd_colsum_workspace_bytes(rows, c)) {
    set_error("gcl_prelu_bwd_colsum_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    cudaMemsetAsync(dslope, 0, sizeof(float), s);
    cudaMemsetAsync(dbias, 0, sizeof(float) * c, s);
    return GCL_OK;
  }
  LnPlan pl = ln_plan(rows);
  float* part = static_cast<float*>(workspace);
  const bool v4 = (c & 3) == 0 && c <= 1024 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) |
                                                   reinterpret_cast<uintptr_t>(dx)) & 15u) == 0;
  if (v4)
    prelu_bwd_colsum_v4_kernel<<<pl.nblk, 256, 0, s>>>(reinterpret_cast<const float4*>(dy),
                                                       reinterpret_cast<const float4*>(x), slope,
                                                       reinterpret_cast<float4*>(dx), part, rows, (int)c, pl.rows_per_block);
  else
    prelu_bwd_colsum_kernel<<<pl.nblk, dim3(32, 8), 0, s>>>(dy, x, slope, dx, part, rows, (int)c, pl.rows_per_block);
  GCL_CHECK_LAUNCH("gcl_prelu_bwd_colsum_f32");
  reduce_cols_kernel<<<(unsigned)ceil_div(c + 1, 32), dim3(32, 32), 0, s>>>(part, pl.nblk, (int)c + 1, dbias, dslope, (int)c);
  GCL_CHECK_LAUNCH("gcl_prelu_bwd_colsum_f32(reduce)");
  return GCL_OK;
}

extern "C" int gcl_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                     float* rstd, int64_t rows, int64_t c, float eps, void* stream) {
  GCL_CHECK_ARG(x && y && rows >= 0 && c > 0, "gcl_layernorm_fwd_f32: bad argument");
  GCL_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "gcl_layernorm_fwd_f32: gamma and beta go together");
  if (c > 1024) {
    set_error("gcl_layernorm_fwd_f32: channels %lld > 1024 unsupported", (long long)c);
    return GCL_ERR_UNSUPPORTED;
  }
  if (rows == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                    reinterpret_cast<uintptr_t>(beta)) & 15u) == 0;
  if ((c & 3) == 0 && c <= 128 && al) {
    const LnPlan pl = ln_plan(rows);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    if (c <= 32) layernorm_fwd_v4_kernel<8><<<pl.nblk, 256, 0, s>>>(x4, gamma, beta, y4, mean, rstd, rows, (int)c, eps, pl.rows_per_block);
    else if (c <= 64) layernorm_fwd_v4_kernel<16><<<pl.nblk, 256, 0, s>>>(x4, gamma, beta, y4, mean, rstd, rows, (int)c, eps, pl.rows_per_block);
    else layernorm_fwd_v4_kernel<32><<<pl.nblk, 256, 0, s>>>(x4, gamma, beta, y4, mean, rstd, rows, (int)c, eps, pl.rows_per_block);
    GCL_CHECK_LAUNCH("gcl_layernorm_fwd_f32(v4)");
    return GCL_OK;
  }
  const unsigned grid = (unsigned)ceil_div(rows * 32, 256);
  if (c <= 128) layernorm_fwd_kernel<4><<<grid, 256, 0, s>>>(x, gamma, beta, y, mean, rstd, rows, (int)c, eps);
  else if (c <= 256) layernorm_fwd_kernel<8><<<grid, 256, 0, s>>>(x, gamma, beta, y, mean, rstd, rows, (int)c, eps);
  else layernorm_fwd_kernel<32><<<grid, 256, 0, s>>>(x, gamma, beta, y, mean, rstd, rows, (int)c, eps);
  GCL_CHECK_LAUNCH("gcl_layernorm_fwd_f32");
  return GCL_OK;
}

extern "C" size_t gcl_layernorm_bwd_workspace_bytes(int64_t rows, int64_t c) {
  if (rows < 0 || c <= 0) return 0;
  return (size_t)ln_plan(rows).nblk * 2 * (size_t)c * sizeof(float) + 256;
}

extern "C" int gcl_layernorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* mean,
                                     const float* rstd, float* dx, float* dgamma, float* dbeta, int64_t rows,
                                     int64_t c, void* workspace, size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(dy && x && mean && rstd && dx && rows >= 0 && c > 0, "gcl_layernorm_bwd_f32: bad argument");
  GCL_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "gcl_layernorm_bwd_f32: dgamma and dbeta go together");
  if (c > 1024) {
    set_error("gcl_layernorm_bwd_f32: channels %lld > 1024 unsupported", (long long)c);
    return GCL_ERR_UNSUPPORTED;
  }
  const bool want_params = dgamma != nullptr;
  if (want_params && (!workspace || workspace_bytes < gcl_layernorm_bwd_workspace_bytes(rows, c))) {
    set_error("gcl_layernorm_bwd_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    if (want_params) {
      cudaMemsetAsync(dgamma, 0, sizeof(float) * c, s);
      cudaMemsetAsync(dbeta, 0, sizeof(float) * c, s);
    }
    return GCL_OK;
  }
  LnPlan pl = ln_plan(rows);
  float* part = want_params ? static_cast<float*>(workspace) : nullptr;
  const size_t smem = (size_t)8 * 2 * c * sizeof(float);
  const int C = (int)c;
  const bool al = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
                    reinterpret_cast<uintptr_t>(gamma)) & 15u) == 0;
  if ((c & 3) == 0 && c <= 128 && al) {
    const float4* dy4 = reinterpret_cast<const float4*>(dy);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* dx4 = reinterpret_cast<float4*>(dx);
    if (c <= 32) layernorm_bwd_v4_kernel<8><<<pl.nblk, 256, 0, s>>>(dy4, x4, gamma, mean, rstd, dx4, part, rows, C, pl.rows_per_block);
    else if (c <= 64) layernorm_bwd_v4_kernel<16><<<pl.nblk, 256, 0, s>>>(dy4, x4, gamma, mean, rstd, dx4, part, rows, C, pl.rows_per_block);
    else layernorm_bwd_v4_kernel<32><<<pl.nblk, 256, 0, s>>>(dy4, x4, gamma, mean, rstd, dx4, part, rows, C, pl.rows_per_block);
  } else if (c <= 128)
    layernorm_bwd_kernel<4><<<pl.nblk, 256, smem, s>>>(dy, x, gamma, mean, rstd, dx, part, rows, C, pl.rows_per_block);
  else if (c <= 256)
    layernorm_bwd_kernel<8><<<pl.nblk, 256, smem, s>>>(dy, x, gamma, mean, rstd, dx, part, rows, C, pl.rows_per_block);
  else {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(layernorm_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
      if (e != cudaSuccess) return fail_cuda(e, "gcl_layernorm_bwd_f32(smem attr)");
    }
    layernorm_bwd_kernel<32><<<pl.nblk, 256, smem, s>>>(dy, x, gamma, mean, rstd, dx, part, rows, C,
                                                        pl.rows_per_block);
  }
  GCL_CHECK_LAUNCH("gcl_layernorm_bwd_f32");
  if (want_params) {
    reduce_cols_kernel<<<(unsigned)ceil_div(2 * c, 32), dim3(32, 32), 0, s>>>(part, pl.nblk, 2 * C, dgamma, dbeta, C);
    GCL_CHECK_LAUNCH("gcl_layernorm_bwd_f32(reduce)");
  }
  return GCL_OK;
}

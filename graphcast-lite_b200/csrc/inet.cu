// Pieces the InteractionNet processor (/root/reference/src/models.py:166-285) and the v2
// configs need on top of the message-passing kernels, fp32, deterministic:
//   gcl_act_fwd/bwd_f32              ReLU / SiLU ("swish")                       models.py:154-163 (_get_activation)
//   gcl_layernorm_graph_fwd/bwd_f32  torch_geometric LayerNorm(mode="graph")     models.py:201-203 (edge_norm)
//                                    y = (x - mean) / (std(unbiased=False) + eps) * gamma + beta, statistics over ALL
//                                    elements of a sample
//   gcl_add_f32                      residual connections                        models.py:226-227
#include "common.cuh"

namespace gcl {
namespace {

constexpr int kT = 256;
inline int blocks_for(int64_t n, int cap = 8 * kNumSMs) {
  int64_t b = ceil_div(n, kT);
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ float act_f(float x, int kind) {
  if (kind == GCL_ACT_RELU) return x > 0.f ? x : 0.f;
  return x / (1.f + __expf(-x));                       // SiLU
}
__device__ __forceinline__ float act_df(float x, int kind) {
  if (kind == GCL_ACT_RELU) return x > 0.f ? 1.f : 0.f;
  const float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}

__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int kind) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) y[i] = act_f(x[i], kind);
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx,
                               int64_t n, int kind) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) dx[i] = dy[i] * act_df(x[i], kind);
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) y[i] = a[i] + b[i];
}

constexpr int kLnBlocks = 256;        // partial sums per sample

__device__ __forceinline__ double block_sum(double v, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int k = 0; k < kT / 32; ++k) t += sm[k];
  __syncthreads();
  return t;                            // valid in thread 0
}

// part[b][blk][2] = (sum a, sum a * c) over the block's element range, float64.
//   forward:  a = x, c = x                          -> sum x, sum x^2
//   backward: a = dy * gamma, c = (x - mean) * inv  -> sum g, sum g * xhat
__global__ void lng_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                   const float* __restrict__ gamma, const float* __restrict__ stats, double* __restrict__ part,
                                   int64_t n, int C) {
  __shared__ double sm[kT / 32];
  const int64_t b = blockIdx.y;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x, beg = blockIdx.x * per, end = min(n, beg + per);
  const float* xb = x + b * n;
  double s0 = 0.0, s1 = 0.0;
  if (dy) {
    const float mean = stats[3 * b], inv = stats[3 * b + 1];
    const float* db = dy + b * n;
    for (int64_t i = beg + threadIdx.x; i < end; i += kT) {
      const float g = db[i] * (gamma ? __ldg(gamma + (int)(i % C)) : 1.f);
      s0 += g;
      s1 += (double)g * ((xb[i] - mean) * inv);
    }
  } else {
    for (int64_t i = beg + threadIdx.x; i < end; i += kT) {
      const float v = xb[i];
      s0 += v;
      s1 += (double)v * v;
    }
  }
  const double t0 = block_sum(s0, sm), t1 = block_sum(s1, sm);
  if (threadIdx.x == 0) {
    part[(b * gridDim.x + blockIdx.x) * 2] = t0;
    part[(b * gridDim.x + blockIdx.x) * 2 + 1] = t1;
  }
}

// forward: stats[b] = (mean, 1 / (std + eps), std); backward: coef[b] = (mean g, mean g xhat)
__global__ void lng_finish_kernel(const double* __restrict__ part, int nblk, int64_t n, float eps, float* __restrict__ out,
                                  int backward) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < nblk; ++k) {
    s0 += part[((int64_t)b * nblk + k) * 2];
    s1 += part[((int64_t)b * nblk + k) * 2 + 1];
  }
  if (backward) {
    out[2 * b] = (float)(s0 / (double)n);
    out[2 * b + 1] = (float)(s1 / (double)n);
  } else {
    const double mean = s0 / (double)n;
    const double var = fmax(s1 / (double)n - mean * mean, 0.0);
    const double sd = sqrt(var);
    out[3 * b] = (float)mean;
    out[3 * b + 1] = (float)(1.0 / (sd + (double)eps));
    out[3 * b + 2] = (float)sd;
  }
}

__global__ void lng_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ stats, float* __restrict__ y, int64_t n, int C, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / n;
    const int c = (int)((i - b * n) % C);
    const float xh = (x[i] - stats[3 * b]) * stats[3 * b + 1];
    y[i] = gamma ? fmaf(xh, __ldg(gamma + c), __ldg(beta + c)) : xh;
  }
}

// dx = inv (g - mean g) - xhat * mean(g xhat) / std;  t = dy * xhat (for d gamma = column sums of t)
__global__ void lng_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                                     const float* __restrict__ stats, const float* __restrict__ coef, float* __restrict__ dx,
                                     float* __restrict__ t_out, int64_t n, int C, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / n;
    const int c = (int)((i - b * n) % C);
    const float mean = stats[3 * b], inv = stats[3 * b + 1], sd = fmaxf(stats[3 * b + 2], 1e-30f);
    const float xh = (x[i] - mean) * inv;
    const float d = dy[i];
    const float g = d * (gamma ? __ldg(gamma + c) : 1.f);
    dx[i] = inv * (g - coef[2 * b]) - xh * (coef[2 * b + 1] / sd);
    if (t_out) t_out[i] = d * xh;
  }
}

}  // namespace
}  // namespace gcl

using namespace gcl;

extern "C" int gcl_act_fwd_f32(const float* x, float* y, int64_t n, int kind, void* stream) {
  GCL_CHECK_ARG(x && y && n >= 0 && (kind == GCL_ACT_RELU || kind == GCL_ACT_SILU), "gcl_act_fwd_f32: bad argument");
  if (n == 0) return GCL_OK;
  act_fwd_kernel<<<blocks_for(n), kT, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, kind);
  GCL_CHECK_LAUNCH("gcl_act_fwd_f32");
  return GCL_OK;
}

extern "C" int gcl_act_bwd_f32(const float* dy, const float* x, float* dx, int64_t n, int kind, void* stream) {
  GCL_CHECK_ARG(dy && x && dx && n >= 0 && (kind == GCL_ACT_RELU || kind == GCL_ACT_SILU), "gcl_act_bwd_f32: bad argument");
  if (n == 0) return GCL_OK;
  act_bwd_kernel<<<blocks_for(n), kT, 0, static_cast<cudaStream_t>(stream)>>>(dy, x, dx, n, kind);
  GCL_CHECK_LAUNCH("gcl_act_bwd_f32");
  return GCL_OK;
}

extern "C" int gcl_add_f32(const float* a, const float* b, float* y, int64_t n, void* stream) {
  GCL_CHECK_ARG(a && b && y && n >= 0, "gcl_add_f32: bad argument");
  if (n == 0) return GCL_OK;
  add_kernel<<<blocks_for(n), kT, 0, static_cast<cudaStream_t>(stream)>>>(a, b, y, n);
  GCL_CHECK_LAUNCH("gcl_add_f32");
  return GCL_OK;
}

extern "C" size_t gcl_layernorm_graph_workspace_bytes(int64_t batch) {
  return batch < 0 ? 0 : (size_t)batch * kLnBlocks * 2 * sizeof(double) + 256;
}

extern "C" int gcl_layernorm_graph_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* stats,
                                           int64_t batch, int64_t elems_per_sample, int64_t c, float eps, void* workspace,
                                           size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(x && y && stats && workspace && batch >= 0 && batch <= 65535 && elems_per_sample > 0 && c > 0 &&
                    elems_per_sample % c == 0 && (!gamma == !beta),
                "gcl_layernorm_graph_fwd_f32: bad argument");
  if (workspace_bytes < gcl_layernorm_graph_workspace_bytes(batch)) {
    set_error("gcl_layernorm_graph_fwd_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  if (batch == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(workspace);
  lng_partial_kernel<<<dim3(kLnBlocks, (unsigned)batch), kT, 0, s>>>(x, nullptr, nullptr, nullptr, part, elems_per_sample, (int)c);
  GCL_CHECK_LAUNCH("gcl_layernorm_graph_fwd_f32(partial)");
  lng_finish_kernel<<<(unsigned)batch, 32, 0, s>>>(part, kLnBlocks, elems_per_sample, eps, stats, 0);
  GCL_CHECK_LAUNCH("gcl_layernorm_graph_fwd_f32(finish)");
  const int64_t total = batch * elems_per_sample;
  lng_apply_kernel<<<blocks_for(total, 16 * kNumSMs), kT, 0, s>>>(x, gamma, beta, stats, y, elems_per_sample, (int)c, total);
  GCL_CHECK_LAUNCH("gcl_layernorm_graph_fwd_f32(apply)");
  return GCL_OK;
}

extern "C" int gcl_layernorm_graph_bwd_f32(const float* dy, const float* x, const float* gamma, const float* stats,
                                           float* dx, float* t_out, float* coef, int64_t batch, int64_t elems_per_sample,
                                           int64_t c, void* workspace, size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(dy && x && stats && dx && coef && workspace && batch >= 0 && batch <= 65535 && elems_per_sample > 0 &&
                    c > 0 && elems_per_sample % c == 0,
                "gcl_layernorm_graph_bwd_f32: bad argument");
  if (workspace_bytes < gcl_layernorm_graph_workspace_bytes(batch)) {
    set_error("gcl_layernorm_graph_bwd_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  if (batch == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(workspace);
  lng_partial_kernel<<<dim3(kLnBlocks, (unsigned)batch), kT, 0, s>>>(x, dy, gamma, stats, part, elems_per_sample, (int)c);
  GCL_CHECK_LAUNCH("gcl_layernorm_graph_bwd_f32(partial)");
  lng_finish_kernel<<<(unsigned)batch, 32, 0, s>>>(part, kLnBlocks, elems_per_sample, 0.f, coef, 1);
  GCL_CHECK_LAUNCH("gcl_layernorm_graph_bwd_f32(finish)");
  const int64_t total = batch * elems_per_sample;
  lng_bwd_apply_kernel<<<blocks_for(total, 16 * kNumSMs), kT, 0, s>>>(x, dy, gamma, stats, coef, dx, t_out,
                                                                      elems_per_sample, (int)c, total);
  GCL_CHECK_LAUNCH("gcl_layernorm_graph_bwd_f32(apply)");
  return GCL_OK;
}

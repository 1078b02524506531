// Glue kernels of one forecast / training step (fp32, deterministic):
//   gcl_assemble_input_f32  WeatherPrediction._preprocess_input   /root/reference/src/models.py:776-806
//   gcl_wmse_f32            residual add + latitude-weighted MSE  /root/reference/src/train.py:85-102,203-213
//   gcl_adam_f32            torch.optim.Adam step                 /root/reference/src/main.py:212, train.py:233
#include "common.cuh"

namespace gcl {
namespace {

constexpr int kT = 256;
constexpr int kMaxBlocks = 8 * kNumSMs;

__global__ void assemble_kernel(const float* __restrict__ x, const float* __restrict__ gs,
                                const float* __restrict__ ms, float* __restrict__ out, int64_t B, int64_t G,
                                int64_t M, int TF, int S) {
  const int W = TF + S;
  const int64_t total = B * (G + M) * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int c = (int)(idx % W);
    const int64_t rn = idx / W;
    const int64_t n = rn % (G + M), b = rn / (G + M);
    float v;
    if (n < G) v = c < TF ? x[(b * G + n) * TF + c] : __ldg(gs + n * S + (c - TF));
    else v = c < TF ? 0.f : __ldg(ms + (n - G) * S + (c - TF));
    out[idx] = v;
  }
}

// Each block owns a contiguous element range [beg, end) of the [B, G, C] problem.
__global__ void wmse_kernel(const float* __restrict__ delta, const float* __restrict__ x_last, int64_t xl_stride,
                            const float* __restrict__ y, int64_t y_stride, const float* __restrict__ lat_w,
                            float* __restrict__ out_state, float* __restrict__ d_delta, float* __restrict__ part,
                            float gscale, int64_t G, int C, int64_t total, int64_t per_block) {
  __shared__ float sm[kT / 32];
  const int64_t beg = (int64_t)blockIdx.x * per_block, end = min(total, beg + per_block);
  float s = 0.f;
  for (int64_t idx = beg + threadIdx.x; idx < end; idx += kT) {
    const int c = (int)(idx % C);
    const int64_t bg = idx / C;
    const int64_t g = bg % G;
    float o = delta[idx];
    if (x_last) o += x_last[bg * xl_stride + c];
    const float diff = o - y[bg * y_stride + c];
    const float w = lat_w ? __ldg(lat_w + g) : 1.f;
    if (out_state) out_state[idx] = o;
    if (d_delta) d_delta[idx] = gscale * w * diff;
    s = fmaf(w * diff, diff, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kT / 32; ++k) t += sm[k];
    part[blockIdx.x] = t;
  }
}

__global__ void wmse_finish_kernel(const float* __restrict__ part, int n, float inv_wsum, float* __restrict__ loss,
                                   int accumulate) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += part[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) {
    const float v = s * inv_wsum;
    *loss = accumulate ? *loss + v : v;
  }
}

__global__ void adam_tick_kernel(int32_t* step) { *step += 1; }

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float gscale,
                            const int32_t* __restrict__ step) {
  const float t = (float)*step;
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

int blocks_for(int64_t n) {
  int64_t b = ceil_div(n, kT);
  return (int)(b < 1 ? 1 : (b > kMaxBlocks ? kMaxBlocks : b));
}

}  // namespace
}  // namespace gcl

using namespace gcl;

extern "C" int gcl_assemble_input_f32(const float* x, const float* grid_static, const float* mesh_static,
                                      float* enc_in, int64_t batch, int64_t n_grid, int64_t n_mesh, int64_t tf,
                                      int64_t s_dim, void* stream) {
  GCL_CHECK_ARG(x && grid_static && mesh_static && enc_in, "gcl_assemble_input_f32: null pointer argument");
  GCL_CHECK_ARG(batch >= 0 && n_grid >= 0 && n_mesh >= 0 && tf >= 0 && s_dim >= 0 && tf + s_dim > 0,
                "gcl_assemble_input_f32: bad sizes");
  const int64_t total = batch * (n_grid + n_mesh) * (tf + s_dim);
  if (total == 0) return GCL_OK;
  assemble_kernel<<<blocks_for(total), kT, 0, static_cast<cudaStream_t>(stream)>>>(
      x, grid_static, mesh_static, enc_in, batch, n_grid, n_mesh, (int)tf, (int)s_dim);
  GCL_CHECK_LAUNCH("gcl_assemble_input_f32");
  return GCL_OK;
}

extern "C" size_t gcl_wmse_workspace_bytes(int64_t batch, int64_t n_grid, int64_t c) {
  (void)batch; (void)n_grid; (void)c;
  return (size_t)kMaxBlocks * sizeof(float) + 256;
}

extern "C" int gcl_wmse_f32(const float* delta, const float* x_last, int64_t xl_stride, const float* y,
                            int64_t y_stride, const float* lat_w, float inv_wsum, float* out_state, float* d_delta,
                            float* loss_out, int accumulate, float scale, int64_t batch, int64_t n_grid, int64_t c,
                            void* workspace, size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(delta && y && loss_out && workspace, "gcl_wmse_f32: null pointer argument");
  GCL_CHECK_ARG(batch > 0 && n_grid > 0 && c > 0, "gcl_wmse_f32: bad sizes");
  if (workspace_bytes < gcl_wmse_workspace_bytes(batch, n_grid, c)) {
    set_error("gcl_wmse_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = batch * n_grid * c;
  const int nblk = blocks_for(total);
  const int64_t per_block = ceil_div(total, nblk);
  float* part = static_cast<float*>(workspace);
  wmse_kernel<<<nblk, kT, 0, s>>>(delta, x_last, xl_stride, y, y_stride, lat_w, out_state, d_delta, part,
                                  scale * 2.f * inv_wsum, n_grid, (int)c, total, per_block);
  GCL_CHECK_LAUNCH("gcl_wmse_f32");
  wmse_finish_kernel<<<1, 32, 0, s>>>(part, nblk, inv_wsum * scale, loss_out, accumulate);
  GCL_CHECK_LAUNCH("gcl_wmse_f32(finish)");
  return GCL_OK;
}

extern "C" int gcl_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                            float beta1, float beta2, float eps, float grad_scale, int32_t* step_count,
                            void* stream) {
  GCL_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step_count && n >= 0, "gcl_adam_f32: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  adam_tick_kernel<<<1, 1, 0, s>>>(step_count);
  GCL_CHECK_LAUNCH("gcl_adam_f32(tick)");
  if (n == 0) return GCL_OK;
  adam_kernel<<<blocks_for(n), kT, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, grad_scale,
                                           step_count);
  GCL_CHECK_LAUNCH("gcl_adam_f32");
  return GCL_OK;
}

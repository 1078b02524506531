// Glue kernels of one forecast / training step (fp32, deterministic):
//   gcl_assemble_input_f32  WeatherPrediction._preprocess_input   /root/reference/src/models.py:776-806
//   gcl_wmse_f32            residual add + latitude-weighted MSE  /root/reference/src/train.py:85-102,203-213
//   gcl_adam_f32            torch.optim.Adam step                 /root/reference/src/main.py:212, train.py:233
#include "common.cuh"

namespace gcl {
namespace {

constexpr int kT = 256;
constexpr int kMaxBlocks = 8 * kNumSMs;

// V output floats per thread (V = 4 when the row width W is a multiple of 4 and `out` is 16-byte aligned: one
// 16-byte store), flat grid-stride over all B * (G + M) * W / V words: consecutive threads write consecutive words
// and read (nearly) consecutive source floats.  (The first version walked a warp per 44-float row: two partly
// filled 32-lane passes per row, 0.27 of the HBM roofline; this one streams.)
// W >= TF + S is the output row width; columns past TF + S are zero (alignment padding)
template <int V>
__global__ void assemble_kernel(const float* __restrict__ x, const float* __restrict__ gs,
                                const float* __restrict__ ms, float* __restrict__ out, int64_t B, int64_t G,
                                int64_t M, int TF, int S, int W) {
  const int wpr = W / V;                                  // words per row
  const int64_t N = G + M, total = B * N * wpr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t rn = i / wpr;
    const int c0 = (int)(i - rn * wpr) * V;
    const int64_t b = rn / N, n = rn - b * N;
    const bool grid_row = n < G;
    const float* xr = x + (b * G + n) * TF;
    const float* sr = grid_row ? gs + n * S : ms + (n - G) * S;
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int c = c0 + k;
      v[k] = c < TF ? (grid_row ? __ldg(xr + c) : 0.f) : (c < TF + S ? __ldg(sr + (c - TF)) : 0.f);
    }
    if (V == 4) *reinterpret_cast<float4*>(out + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
    else out[i] = v[0];
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

// One autoregressive training step behind the model (train.py:201-227), one pass over [B, G, C]:
//   out = (residual ? state[.., obs-1, :] : 0) + delta;  loss part = sum w (out - y)^2, w = node_w[g] chan_w[c];
//   g_loss = gscale w (out - y);  new_state = [state[.., 1:, :], out'] with out'[c] = static ? x_last : forcing ? y : out
__global__ void ar_step_kernel(const float* __restrict__ delta, const float* __restrict__ state,
                               const float* __restrict__ y, int64_t y_stride, const float* __restrict__ node_w,
                               const float* __restrict__ chan_w, const int32_t* __restrict__ carry, int residual,
                               float* __restrict__ new_state, float* __restrict__ g_loss, float* __restrict__ part,
                               float gscale, int64_t G, int obs, int C, int64_t total, int64_t per_block) {
  __shared__ float sm[kT / 32];
  const int64_t beg = (int64_t)blockIdx.x * per_block, end = min(total, beg + per_block);
  float s = 0.f;
  for (int64_t idx = beg + threadIdx.x; idx < end; idx += kT) {
    const int c = (int)(idx % C);
    const int64_t bg = idx / C;
    const int64_t g = bg % G;
    const float* st = state + bg * (int64_t)obs * C + c;
    const float xl = st[(int64_t)(obs - 1) * C];
    const float yv = y ? y[bg * y_stride + c] : 0.f;
    const float o = delta[idx] + (residual ? xl : 0.f);
    const float diff = o - yv;
    const float w = (node_w ? __ldg(node_w + g) : 1.f) * (chan_w ? __ldg(chan_w + c) : 1.f);
    if (g_loss) g_loss[idx] = gscale * w * diff;
    s = fmaf(w * diff, diff, s);
    if (new_state) {
      float* ns = new_state + bg * (int64_t)obs * C + c;
      for (int t = 0; t + 1 < obs; ++t) ns[(int64_t)t * C] = st[(int64_t)(t + 1) * C];
      const int cr = carry ? __ldg(carry + c) : 0;
      ns[(int64_t)(obs - 1) * C] = cr == 1 ? xl : (cr == 2 ? yv : o);
    }
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

// backward of ar_step: d_delta = g_loss dloss + (carry == 0) d_new[obs-1];
//   d_state[t] = (t >= 1 ? d_new[t-1] : 0) + (t == obs-1 ? residual d_delta + (carry == 1) d_new[obs-1] : 0)
__global__ void ar_step_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ dloss,
                                   const float* __restrict__ d_new, const int32_t* __restrict__ carry, int residual,
                                   float* __restrict__ d_delta, float* __restrict__ d_state, int obs, int C,
                                   int64_t total) {
  const float dl = dloss ? __ldg(dloss) : 1.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int c = (int)(idx % C);
    const int64_t bg = idx / C;
    const int cr = carry ? __ldg(carry + c) : 0;
    const float* dn = d_new ? d_new + bg * (int64_t)obs * C + c : nullptr;
    const float dlast = dn ? dn[(int64_t)(obs - 1) * C] : 0.f;
    const float dd = g_loss[idx] * dl + (cr == 0 ? dlast : 0.f);
    d_delta[idx] = dd;
    if (d_state) {
      float* ds = d_state + bg * (int64_t)obs * C + c;
      for (int t = 0; t < obs; ++t) {
        float v = (t >= 1 && dn) ? dn[(int64_t)(t - 1) * C] : 0.f;
        if (t == obs - 1) v += (residual ? dd : 0.f) + (cr == 1 ? dlast : 0.f);
        ds[(int64_t)t * C] = v;
      }
    }
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

// out[b] = [a[b]; b_[b]] along the node axis (CONCAT) or the reverse (split).  One element = VEC floats of a row
// segment; segments are contiguous per sample, so each thread walks whole 16-byte words.
template <typename T, bool CONCAT>
__global__ void rows_cat_kernel(const T* __restrict__ a_c, const T* __restrict__ b_c, T* __restrict__ x_m,
                                T* __restrict__ a_m, T* __restrict__ b_m, const T* __restrict__ x_c, int64_t B,
                                int64_t ea, int64_t eb) {
  // ea / eb: elements (of T) per sample in the a / b part
  const int64_t per = ea + eb, total = B * per;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / per, r = i - b * per;
    const bool first = r < ea;
    const int64_t part_idx = first ? b * ea + r : b * eb + (r - ea);
    if (CONCAT) {
      const T* src = first ? a_c : b_c;
      T v{};
      if (src) v = src[part_idx];
      x_m[i] = v;
    } else {
      T* dst = first ? a_m : b_m;
      if (dst) dst[part_idx] = x_c[i];
    }
  }
}

// dst[b][dst_row0 + r] = src[b][src_row0 + r] for r < rows: one block of node rows of every sample, between tensors
// with different node counts (the mesh rows of [B, G + M, C] <-> a contiguous [B, M, C])
template <typename T>
__global__ void rows_block_copy_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t B, int64_t e_block,
                                       int64_t src_off, int64_t src_per, int64_t dst_off, int64_t dst_per) {
  const int64_t total = B * e_block;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / e_block, r = i - b * e_block;
    dst[b * dst_per + dst_off + r] = src[b * src_per + src_off + r];
  }
}

// y[r, 0:c_out] = x[r, 0:min(c_in, c_out)], zero beyond: the slice that drops a layer's alignment padding, and its
// backward (zero padding) -- one pass instead of torch's zero fill + strided copy
__global__ void resize_channels_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int c_in,
                                       int c_out, int vec) {
  // four consecutive output floats per thread (one 64-bit division, one 16-byte store when y is 16-byte aligned)
  const int64_t total = rows * c_out, groups = (total + 3) / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gi < groups; gi += stride) {
    const int64_t i = gi * 4;
    int64_t r = i / c_out;
    int c = (int)(i - r * c_out);
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = (i + k < total && c < c_in) ? __ldg(x + r * c_in + c) : 0.f;
      if (++c == c_out) { c = 0; ++r; }
    }
    if (vec && i + 3 < total) {
      *reinterpret_cast<float4*>(y + i) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i + k < total) y[i + k] = v[k];
    }
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
                                      int64_t s_dim, int64_t out_width, void* stream) {
  GCL_CHECK_ARG(x && grid_static && mesh_static && enc_in, "gcl_assemble_input_f32: null pointer argument");
  GCL_CHECK_ARG(batch >= 0 && n_grid >= 0 && n_mesh >= 0 && tf >= 0 && s_dim >= 0 && tf + s_dim > 0,
                "gcl_assemble_input_f32: bad sizes");
  GCL_CHECK_ARG(out_width >= tf + s_dim, "gcl_assemble_input_f32: out_width %lld < tf + s = %lld", (long long)out_width,
                (long long)(tf + s_dim));
  const int64_t total = batch * (n_grid + n_mesh) * out_width;
  if (total == 0) return GCL_OK;
  if (out_width % 4 == 0 && (reinterpret_cast<uintptr_t>(enc_in) & 15u) == 0)
    assemble_kernel<4><<<blocks_for(total / 4), kT, 0, static_cast<cudaStream_t>(stream)>>>(
        x, grid_static, mesh_static, enc_in, batch, n_grid, n_mesh, (int)tf, (int)s_dim, (int)out_width);
  else
    assemble_kernel<1><<<blocks_for(total), kT, 0, static_cast<cudaStream_t>(stream)>>>(
        x, grid_static, mesh_static, enc_in, batch, n_grid, n_mesh, (int)tf, (int)s_dim, (int)out_width);
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

namespace {
template <bool CONCAT>
int rows_cat(const float* a, const float* b_, float* x, int64_t batch, int64_t na, int64_t nb, int64_t c, void* stream) {
  const int64_t ea = na * c, eb = nb * c;
  if (batch == 0 || ea + eb == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool v4 = (ea % 4 == 0) && (eb % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b_) | reinterpret_cast<uintptr_t>(x)) & 15u) == 0;
  if (v4) {
    const int64_t total = batch * (ea + eb) / 4;
    gcl::rows_cat_kernel<float4, CONCAT><<<gcl::blocks_for(total), gcl::kT, 0, s>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b_), reinterpret_cast<float4*>(x),
        reinterpret_cast<float4*>(const_cast<float*>(a)), reinterpret_cast<float4*>(const_cast<float*>(b_)),
        reinterpret_cast<const float4*>(x), batch, ea / 4, eb / 4);
  } else {
    gcl::rows_cat_kernel<float, CONCAT><<<gcl::blocks_for(batch * (ea + eb)), gcl::kT, 0, s>>>(
        a, b_, x, const_cast<float*>(a), const_cast<float*>(b_), x, batch, ea, eb);
  }
  return GCL_OK;
}
}  // namespace

extern "C" int gcl_rows_concat_f32(const float* a, const float* b_, float* out, int64_t batch, int64_t na, int64_t nb,
                                   int64_t c, void* stream) {
  GCL_CHECK_ARG(out && batch >= 0 && na >= 0 && nb >= 0 && c > 0, "gcl_rows_concat_f32: bad argument");
  const int rc = rows_cat<true>(a, b_, out, batch, na, nb, c, stream);
  GCL_CHECK_LAUNCH("gcl_rows_concat_f32");
  return rc;
}

extern "C" int gcl_rows_block_copy_f32(const float* src, float* dst, int64_t batch, int64_t rows, int64_t c,
                                       int64_t src_row0, int64_t src_rows, int64_t dst_row0, int64_t dst_rows,
                                       void* stream) {
  GCL_CHECK_ARG(src && dst && batch >= 0 && rows >= 0 && c > 0, "gcl_rows_block_copy_f32: bad argument");
  GCL_CHECK_ARG(src_row0 >= 0 && dst_row0 >= 0 && src_row0 + rows <= src_rows && dst_row0 + rows <= dst_rows,
                "gcl_rows_block_copy_f32: row block out of range");
  if (batch == 0 || rows == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t eb = rows * c, so = src_row0 * c, sp = src_rows * c, d_o = dst_row0 * c, dp = dst_rows * c;
  const bool v4 = ((eb | so | sp | d_o | dp) % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
  if (v4)
    gcl::rows_block_copy_kernel<float4><<<gcl::blocks_for(batch * eb / 4), gcl::kT, 0, s>>>(
        reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), batch, eb / 4, so / 4, sp / 4, d_o / 4, dp / 4);
  else
    gcl::rows_block_copy_kernel<float><<<gcl::blocks_for(batch * eb), gcl::kT, 0, s>>>(src, dst, batch, eb, so, sp, d_o, dp);
  GCL_CHECK_LAUNCH("gcl_rows_block_copy_f32");
  return GCL_OK;
}

extern "C" int gcl_rows_split_f32(const float* x, float* a, float* b_, int64_t batch, int64_t na, int64_t nb, int64_t c,
                                  void* stream) {
  GCL_CHECK_ARG(x && batch >= 0 && na >= 0 && nb >= 0 && c > 0, "gcl_rows_split_f32: bad argument");
  const int rc = rows_cat<false>(a, b_, const_cast<float*>(x), batch, na, nb, c, stream);
  GCL_CHECK_LAUNCH("gcl_rows_split_f32");
  return rc;
}

extern "C" int gcl_ar_step_f32(const float* delta, const float* state, const float* y, int64_t y_stride,
                               const float* node_w, const float* chan_w, const int32_t* carry, int residual,
                               float inv_wsum, float scale, float* new_state, float* g_loss, float* loss_out,
                               int accumulate, int64_t batch, int64_t n_grid, int64_t obs, int64_t c, void* workspace,
                               size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(delta && state && loss_out && workspace, "gcl_ar_step_f32: null pointer argument");
  GCL_CHECK_ARG(batch >= 0 && n_grid > 0 && c > 0 && obs > 0 && y_stride >= c, "gcl_ar_step_f32: bad sizes");
  GCL_CHECK_ARG(new_state != state, "gcl_ar_step_f32: new_state must not alias state");
  if (workspace_bytes < gcl_wmse_workspace_bytes(batch, n_grid, c)) {
    set_error("gcl_ar_step_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = batch * n_grid * c;
  const int nblk = blocks_for(total);
  const int64_t per_block = ceil_div(total, nblk);
  float* part = static_cast<float*>(workspace);
  ar_step_kernel<<<nblk, kT, 0, s>>>(delta, state, y, y_stride, node_w, chan_w, carry, residual, new_state, g_loss, part,
                                     2.f * scale * inv_wsum, n_grid, (int)obs, (int)c, total, per_block);
  GCL_CHECK_LAUNCH("gcl_ar_step_f32");
  wmse_finish_kernel<<<1, 32, 0, s>>>(part, nblk, inv_wsum * scale, loss_out, accumulate);
  GCL_CHECK_LAUNCH("gcl_ar_step_f32(finish)");
  return GCL_OK;
}

extern "C" int gcl_ar_step_bwd_f32(const float* g_loss, const float* dloss, const float* d_new_state,
                                   const int32_t* carry, int residual, float* d_delta, float* d_state, int64_t batch,
                                   int64_t n_grid, int64_t obs, int64_t c, void* stream) {
  GCL_CHECK_ARG(g_loss && d_delta, "gcl_ar_step_bwd_f32: null pointer argument");
  GCL_CHECK_ARG(batch >= 0 && n_grid > 0 && c > 0 && obs > 0, "gcl_ar_step_bwd_f32: bad sizes");
  const int64_t total = batch * n_grid * c;
  if (total == 0) return GCL_OK;
  ar_step_bwd_kernel<<<blocks_for(total), kT, 0, static_cast<cudaStream_t>(stream)>>>(
      g_loss, dloss, d_new_state, carry, residual, d_delta, d_state, (int)obs, (int)c, total);
  GCL_CHECK_LAUNCH("gcl_ar_step_bwd_f32");
  return GCL_OK;
}

extern "C" int gcl_resize_channels_f32(const float* x, float* y, int64_t rows, int64_t c_in, int64_t c_out, void* stream) {
  GCL_CHECK_ARG(x && y && rows >= 0 && c_in > 0 && c_out > 0, "gcl_resize_channels_f32: bad argument");
  if (rows == 0) return GCL_OK;
  resize_channels_kernel<<<blocks_for((rows * c_out + 3) / 4), kT, 0, static_cast<cudaStream_t>(stream)>>>(
      x, y, rows, (int)c_in, (int)c_out, (reinterpret_cast<uintptr_t>(y) & 15u) == 0 ? 1 : 0);
  GCL_CHECK_LAUNCH("gcl_resize_channels_f32");
  return GCL_OK;
}

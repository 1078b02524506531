// Kernels either side of the model, fp32 results, deterministic:
//   gcl_window_assemble      raw (T, lon, lat, F) windows -> normalised, lat-major [B, G, obs*F] / [B, G, pred*F]
//                            TimeseriesChunkDataset.__getitem__  /root/reference/src/data/dataloader_chunked.py:179-223
//   gcl_forecast_metrics_f32 per (sample, column) error / correlation sums of a batch of forecasts
//                            StreamingMetrics.update             /root/reference/scripts/predict.py:53-124
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace gcl {
namespace {

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<double>(double v) { return (float)v; }     // numpy .astype(float32)

// one thread per (sample, node, window step): F contiguous outputs, (x - mean) / std in float32 exactly as the
// reference computes it on the host (dataloader_chunked.py:190-191, 204-207)
template <typename T>
__global__ void window_assemble_kernel(const T* __restrict__ raw, const float* __restrict__ mean,
                                       const float* __restrict__ stdv, float* __restrict__ X, float* __restrict__ Y,
                                       int64_t B, int W, int obs, int64_t nlon, int64_t nlat, int ftot, int F, int flat) {
  const int64_t G = nlon * nlat;
  const int64_t total = B * G * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int pred = W - obs;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
    const int w = (int)(t % W);
    const int64_t bn = t / W, node = bn % G, b = bn / G;
    int64_t src_node = node;                               // flat datasets: node order as stored
    if (!flat) {                                           // (lon, lat) storage -> lat-major node index (:218-221)
      const int64_t la = node / nlon, lo = node - la * nlon;
      src_node = lo * nlat + la;
    }
    const T* r = raw + ((b * W + w) * G + src_node) * ftot;
    float* o = w < obs ? X + (bn * obs + w) * F : Y + (bn * pred + (w - obs)) * F;
    for (int f = 0; f < F; ++f) o[f] = (to_f32<T>(r[f]) - __ldg(mean + f)) / __ldg(stdv + f);
  }
}

constexpr int kMetricSlices = 16;

// part[b][col][slice][7] = sums over the slice's grid nodes of err^2, |err|, yt, yp, yt^2, yp^2, yt*yp (float64).
// block (32 columns, 8 node lanes)
__global__ void metrics_partial_kernel(const float* __restrict__ yt, const float* __restrict__ yp,
                                       double* __restrict__ part, int64_t G, int CP) {
  __shared__ double sm[8][32][7];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int64_t b = blockIdx.y;
  const int slice = blockIdx.z;
  const int64_t g0 = G * slice / kMetricSlices, g1 = G * (slice + 1) / kMetricSlices;
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  if (col < CP) {
    for (int64_t g = g0 + threadIdx.y; g < g1; g += 8) {
      const int64_t i = (b * G + g) * CP + col;
      const float t = yt[i], p = yp[i];
      const float e = p - t;                               // float32 error, as the reference forms it
      a[0] += (double)e * e; a[1] += fabs((double)e); a[2] += t; a[3] += p;
      a[4] += (double)t * t; a[5] += (double)p * p; a[6] += (double)t * p;
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) sm[threadIdx.y][threadIdx.x][k] = a[k];
  __syncthreads();
  if (threadIdx.y == 0 && col < CP) {
    double* o = part + (((b * CP + col) * kMetricSlices) + slice) * 7;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      double s = sm[0][threadIdx.x][k];
#pragma unroll
      for (int l = 1; l < 8; ++l) s += sm[l][threadIdx.x][k];
      o[k] = s;
    }
  }
}

// out[b][col] = {sum err^2, sum |err|, spatial anomaly correlation}: corr = <yt - mt, yp - mp> / (|yt - mt| |yp - mp| + 1e-8)
__global__ void metrics_finish_kernel(const double* __restrict__ part, double* __restrict__ out, int64_t n, int64_t G) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int s = 0; s < kMetricSlices; ++s)
#pragma unroll
    for (int k = 0; k < 7; ++k) a[k] += part[(i * kMetricSlices + s) * 7 + k];
  const double g = (double)G, mt = a[2] / g, mp = a[3] / g;
  const double cov = a[6] - g * mt * mp;
  const double vt = fmax(a[4] - g * mt * mt, 0.0), vp = fmax(a[5] - g * mp * mp, 0.0);
  out[i * 3 + 0] = a[0];
  out[i * 3 + 1] = a[1];
  out[i * 3 + 2] = cov / (sqrt(vt) * sqrt(vp) + 1e-8);
}

}  // namespace
}  // namespace gcl

using namespace gcl;

extern "C" int gcl_window_assemble(const void* raw, int raw_dtype, const float* mean, const float* stdv, float* x_out,
                                   float* y_out, int64_t batch, int64_t window, int64_t obs, int64_t n_lon,
                                   int64_t n_lat, int64_t f_total, int64_t f_used, int flat, void* stream) {
  GCL_CHECK_ARG(raw && mean && stdv && x_out && (y_out || window == obs), "gcl_window_assemble: null pointer argument");
  GCL_CHECK_ARG(batch >= 0 && window > 0 && obs > 0 && obs <= window && n_lon > 0 && n_lat > 0 && f_used > 0 &&
                    f_used <= f_total,
                "gcl_window_assemble: bad sizes");
  const int64_t total = batch * n_lon * n_lat * window;
  if (total == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(total, 256), 16 * kNumSMs);
#define GCL_WA(T)                                                                                                   \
  window_assemble_kernel<T><<<blocks, 256, 0, s>>>(static_cast<const T*>(raw), mean, stdv, x_out, y_out, batch,       \
                                                   (int)window, (int)obs, n_lon, n_lat, (int)f_total, (int)f_used, flat)
  if (raw_dtype == GCL_RAW_F16) GCL_WA(__half);
  else if (raw_dtype == GCL_RAW_F32) GCL_WA(float);
  else if (raw_dtype == GCL_RAW_F64) GCL_WA(double);
  else {
    set_error("gcl_window_assemble: raw_dtype %d is not one of GCL_RAW_F16 / F32 / F64", raw_dtype);
    return GCL_ERR_UNSUPPORTED;
  }
#undef GCL_WA
  GCL_CHECK_LAUNCH("gcl_window_assemble");
  return GCL_OK;
}

extern "C" size_t gcl_forecast_metrics_workspace_bytes(int64_t batch, int64_t cols) {
  if (batch < 0 || cols < 0) return 0;
  return (size_t)batch * cols * kMetricSlices * 7 * sizeof(double) + 256;
}

extern "C" int gcl_forecast_metrics_f32(const float* y_true, const float* y_pred, double* out, int64_t batch,
                                        int64_t n_grid, int64_t cols, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  GCL_CHECK_ARG(y_true && y_pred && out && workspace, "gcl_forecast_metrics_f32: null pointer argument");
  GCL_CHECK_ARG(batch >= 0 && batch <= 65535 && n_grid > 0 && cols > 0, "gcl_forecast_metrics_f32: bad sizes");
  if (workspace_bytes < gcl_forecast_metrics_workspace_bytes(batch, cols)) {
    set_error("gcl_forecast_metrics_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  if (batch == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(workspace);
  metrics_partial_kernel<<<dim3((unsigned)ceil_div(cols, 32), (unsigned)batch, kMetricSlices), dim3(32, 8), 0, s>>>(
      y_true, y_pred, part, n_grid, (int)cols);
  GCL_CHECK_LAUNCH("gcl_forecast_metrics_f32(partial)");
  metrics_finish_kernel<<<(unsigned)ceil_div(batch * cols, 128), 128, 0, s>>>(part, out, batch * cols, n_grid);
  GCL_CHECK_LAUNCH("gcl_forecast_metrics_f32(finish)");
  return GCL_OK;
}

// K7: node-wise dense transforms in fp32 (FFMA, fp32 accumulate) -- the torch.nn.Linear layers of the
// reference MLP (/root/reference/src/models.py:74-98) and the bias-free `lin` inside GCNConv / GATConv
// (PyG: x @ W.T), forward and backward.
//
//   gemm_nn_kernel : C[M,N] = A[M,K] B[K,N] (+bias, PReLU)   forward (B = W^T) and dX = dY W (B = W)
//   gemm_tn_kernel : C[M,N] = sum_r A[r,M] B[r,N]            dW = dY^T X, rows split over CTAs, partial
//                    tiles summed in a fixed order by reduce_partials_kernel (deterministic, no atomics)
//
// fp32 on CUDA cores keeps rel 1e-4 parity with the fp32 reference (a TF32 tensor-core path would not;
// DESIGN.md "dense transform").  Tiles: 128 x {128,64,32} x 16, 256 threads, 8 x {8,4,2} accumulators
// per thread, cp.async double buffering, zero-filled edges.
#include "common.cuh"

namespace gcl {
namespace {

constexpr int kThreads = 256;
constexpr int BK = 16;
constexpr int A_STRIDE = BK + 4;  // floats; rows stay 16 B aligned, neighbouring rows hit different banks

// ---- tile loaders --------------------------------------------------------------------------------
// Copy a [ROWS x COLS] fp32 tile (row-major source with leading dimension ld) into smem with row
// stride SSTRIDE; out-of-range elements become 0.  VEC = 4 -> 16 B cp.async (needs ld % 4 == 0 and
// 16 B aligned base), VEC = 1 -> 4 B cp.async.
template <int ROWS, int COLS, int SSTRIDE, int VEC>
__device__ __forceinline__ void load_tile(float* __restrict__ sm, const float* __restrict__ g, int64_t ld,
                                          int64_t row0, int64_t col0, int64_t n_rows, int64_t n_cols) {
  constexpr int PER_ROW = COLS / VEC;
  constexpr int TOTAL = ROWS * PER_ROW;
#pragma unroll
  for (int it = 0; it < (TOTAL + kThreads - 1) / kThreads; ++it) {
    const int idx = it * kThreads + threadIdx.x;
    if (TOTAL % kThreads != 0 && idx >= TOTAL) break;
    const int r = idx / PER_ROW, c = (idx % PER_ROW) * VEC;
    const int64_t gr = row0 + r, gc = col0 + c;
    const bool ok = (gr < n_rows) && (gc < n_cols);  // VEC = 4: n_cols % 4 == 0, so whole vector is in range
    const float* src = ok ? (g + gr * ld + gc) : g;
    cp_async_zfill<VEC * 4>(sm + r * SSTRIDE + c, src, ok);
  }
}

// T consecutive floats from shared memory (T = 8, 4 or 2; p is 4*T-byte aligned)
template <int T>
__device__ __forceinline__ void lds_frag(float (&r)[T], const float* p) {
  if constexpr (T == 8) {
    const float4 v0 = *reinterpret_cast<const float4*>(p);
    const float4 v1 = *reinterpret_cast<const float4*>(p + 4);
    r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
  } else if constexpr (T == 4) {
    const float4 v0 = *reinterpret_cast<const float4*>(p);
    r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w;
  } else {
    const float2 v0 = *reinterpret_cast<const float2*>(p);
    r[0] = v0.x; r[1] = v0.y;
  }
}

// ---- C = A B (+ epilogue) ------------------------------------------------------------------------
template <int TN, int AVEC, int BVEC>
__global__ void __launch_bounds__(kThreads, 2)
    gemm_nn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int64_t M,
                   int N, int K, const float* __restrict__ bias, const float* __restrict__ prelu_slope,
                   float* __restrict__ z_out, bool vec_out) {
  constexpr int BM = 128, BN = 16 * TN;
  __shared__ __align__(16) float As[2][BM * A_STRIDE];
  __shared__ __align__(16) float Bs[2][BK * BN];
  const int tid = threadIdx.x, tn = tid & 15, tm = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  load_tile<BM, BK, A_STRIDE, AVEC>(As[0], A, K, m0, 0, M, K);
  load_tile<BK, BN, BN, BVEC>(Bs[0], B, N, 0, n0, K, N);
  cp_async_commit();
  for (int kc = 0; kc < nk; ++kc) {
    const int cur = kc & 1;
    if (kc + 1 < nk) {
      load_tile<BM, BK, A_STRIDE, AVEC>(As[cur ^ 1], A, K, m0, (int64_t)(kc + 1) * BK, M, K);
      load_tile<BK, BN, BN, BVEC>(Bs[cur ^ 1], B, N, (int64_t)(kc + 1) * BK, n0, K, N);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* as = As[cur];
    const float* bs = Bs[cur];
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(as + (tm + 16 * i) * A_STRIDE + kk);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float b[TN];
        lds_frag<TN>(b, bs + (kk + k) * BN + tn * TN);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float av = (k == 0) ? a[i].x : (k == 1) ? a[i].y : (k == 2) ? a[i].z : a[i].w;
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
        }
      }
    }
    __syncthreads();
  }

  const float slope = prelu_slope ? __ldg(prelu_slope) : 0.f;
  const int nb = n0 + tn * TN;
  if (vec_out && TN >= 4 && nb + TN <= N) {  // 16 B stores; N % 4 == 0 and 16 B aligned bases
    float bv[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = bias ? __ldg(bias + nb + j) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + tm + 16 * i;
      if (m >= M) continue;
#pragma unroll
      for (int q = 0; q < TN / 4; ++q) {
        float4 v = make_float4(acc[i][4 * q] + bv[4 * q], acc[i][4 * q + 1] + bv[4 * q + 1],
                               acc[i][4 * q + 2] + bv[4 * q + 2], acc[i][4 * q + 3] + bv[4 * q + 3]);
        if (z_out) st4(z_out + m * N + nb + 4 * q, v);
        if (prelu_slope)
          v = make_float4(prelu_f(v.x, slope), prelu_f(v.y, slope), prelu_f(v.z, slope), prelu_f(v.w, slope));
        st4(C + m * N + nb + 4 * q, v);
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + tm + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tn * TN + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + n);
      if (z_out) z_out[m * N + n] = v;
      if (prelu_slope) v = prelu_f(v, slope);
      C[m * N + n] = v;
    }
  }
}

// ---- C[M,N] = sum_r A[r,M] B[r,N] over this CTA's row range (partials) ----------------------------
template <int TM, int TN, int AVEC, int BVEC>
__global__ void __launch_bounds__(kThreads, 2)
    gemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ part,
                   float* __restrict__ part_colsum, int64_t R, int M, int N, int64_t rows_per_split) {
  constexpr int BM = 16 * TM, BN = 16 * TN;
  __shared__ __align__(16) float As[2][BK * BM];
  __shared__ __align__(16) float Bs[2][BK * BN];
  const int tid = threadIdx.x, tn = tid & 15, tm = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.z * BN;
  const int64_t r_beg = (int64_t)blockIdx.x * rows_per_split;
  const int64_t r_end = min(R, r_beg + rows_per_split);

  float acc[TM][TN];
  float csum[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    csum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  }
  const int nk = (int)((r_end - r_beg + BK - 1) / BK);
  if (nk > 0) {
    load_tile<BK, BM, BM, AVEC>(As[0], A, M, r_beg, m0, r_end, M);
    load_tile<BK, BN, BN, BVEC>(Bs[0], B, N, r_beg, n0, r_end, N);
    cp_async_commit();
  }
  for (int kc = 0; kc < nk; ++kc) {
    const int cur = kc & 1;
    if (kc + 1 < nk) {
      load_tile<BK, BM, BM, AVEC>(As[cur ^ 1], A, M, r_beg + (int64_t)(kc + 1) * BK, m0, r_end, M);
      load_tile<BK, BN, BN, BVEC>(Bs[cur ^ 1], B, N, r_beg + (int64_t)(kc + 1) * BK, n0, r_end, N);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* as = As[cur];
    const float* bs = Bs[cur];
#pragma unroll 8
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      lds_frag<TM>(a, as + k * BM + tm * TM);
      lds_frag<TN>(b, bs + k * BN + tn * TN);
      if (part_colsum) {
#pragma unroll
        for (int i = 0; i < TM; ++i) csum[i] += a[i];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  float* p = part + (int64_t)blockIdx.x * M * N;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + tm * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tn * TN + j;
      if (n < N) p[(int64_t)m * N + n] = acc[i][j];
    }
    if (part_colsum && blockIdx.z == 0 && tn == 0) part_colsum[(int64_t)blockIdx.x * M + m] = csum[i];
  }
}

// out[i] = sum_s part[s * n + i] in a fixed order: 32 interleaved partial sums over s (one per threadIdx.y),
// then those 32 in ascending order.  Block (32, 32) per 32 outputs, so even a 64 x 64 dW is spread over 128 CTAs
// (one thread per output summing 148 slices serially took 12-15 us per call, launch-latency sized work).
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n,
                                       int nsplit) {
  __shared__ float sm[32][33];
  const int64_t i = blockIdx.x * 32ll + threadIdx.x;
  float s = 0.f;
  if (i < n)
#pragma unroll 4
    for (int k = threadIdx.y; k < nsplit; k += 32) s += part[(int64_t)k * n + i];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    float t = sm[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < 32; ++y) t += sm[y][threadIdx.x];
    out[i] = t;
  }
}

__global__ void transpose_kernel(const float* __restrict__ W, float* __restrict__ Wt, int rows, int cols) {
  __shared__ float t[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int rr = blockIdx.y * 32 + r;
    if (rr < rows && c < cols) t[r][threadIdx.x] = W[(int64_t)rr * cols + c];
  }
  __syncthreads();
  const int orow = blockIdx.x * 32, ocol = blockIdx.y * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    if (orow + r < cols && ocol < rows) Wt[(int64_t)(orow + r) * rows + ocol] = t[threadIdx.x][r];
  }
}

// ---- column sums ---------------------------------------------------------------------------------
// grid.x = row blocks; block (32, 8).  partial[blk][c] = sum over the block's rows.
__global__ void colsum_partial_kernel(const float* __restrict__ x, float* __restrict__ part, int64_t R, int C,
                                      int64_t rows_per_block) {
  __shared__ float sm[8][33];
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(R, r0 + rows_per_block);
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + threadIdx.x;
    float s = 0.f;
    if (c < C)
      for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += __ldg(x + r * C + c);
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
      part[(int64_t)blockIdx.x * C + c] = t;
    }
    __syncthreads();
  }
}

// 128-bit version (C % 4 == 0, C <= 1024, 16-byte aligned): float4 column groups x 256 / (C / 4) row lanes,
// 4 rows in flight per thread.
__global__ void __launch_bounds__(256)
    colsum_partial_v4_kernel(const float4* __restrict__ x, float* __restrict__ part, int64_t R, int C,
                             int64_t rows_per_block) {
  __shared__ float4 sm[256];
  const int tid = threadIdx.x, ncol = C >> 2, lanes = 256 / ncol;
  const int ci = tid % ncol, rl = tid / ncol;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < lanes) {
    for (int64_t r = r0 + rl; r < r1; r += 4 * lanes) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t rr = r + (int64_t)u * lanes;
        v[u] = rr < r1 ? __ldg(x + rr * ncol + ci) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cs.x += v[u].x; cs.y += v[u].y; cs.z += v[u].z; cs.w += v[u].w;
      }
    }
  }
  sm[tid] = cs;
  __syncthreads();
  if (tid < ncol) {
    float4 t = sm[tid];
    for (int l = 1; l < lanes; ++l) {
      const float4 o = sm[l * ncol + tid];
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    float* dst = part + (int64_t)blockIdx.x * C + 4 * tid;
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int TN>
int launch_nn(const float* A, const float* B, float* C, int64_t M, int N, int K, const float* bias,
              const float* slope, float* z_out, cudaStream_t s) {
  const bool av = (K % 4 == 0) && al16(A), bv = (N % 4 == 0) && al16(B);
  const bool vo = (N % 4 == 0) && al16(C) && (!z_out || al16(z_out));
  dim3 grid((unsigned)ceil_div(M, 128), (unsigned)ceil_div(N, 16 * TN));
  if (av && bv) gemm_nn_kernel<TN, 4, 4><<<grid, kThreads, 0, s>>>(A, B, C, M, N, K, bias, slope, z_out, vo);
  else if (av) gemm_nn_kernel<TN, 4, 1><<<grid, kThreads, 0, s>>>(A, B, C, M, N, K, bias, slope, z_out, vo);
  else if (bv) gemm_nn_kernel<TN, 1, 4><<<grid, kThreads, 0, s>>>(A, B, C, M, N, K, bias, slope, z_out, vo);
  else gemm_nn_kernel<TN, 1, 1><<<grid, kThreads, 0, s>>>(A, B, C, M, N, K, bias, slope, z_out, vo);
  GCL_CHECK_LAUNCH("gemm_nn");
  return GCL_OK;
}

int gemm_nn(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, const float* bias,
            const float* slope, float* z_out, cudaStream_t s) {
  if (M == 0) return GCL_OK;
  if (ceil_div(M, 128) > 2147483647LL || ceil_div(N, 32) > 65535) {
    set_error("gemm_nn: problem too large");
    return GCL_ERR_UNSUPPORTED;
  }
  if (N > 64) return launch_nn<8>(A, B, C, M, (int)N, (int)K, bias, slope, z_out, s);
  if (N > 32) return launch_nn<4>(A, B, C, M, (int)N, (int)K, bias, slope, z_out, s);
  return launch_nn<2>(A, B, C, M, (int)N, (int)K, bias, slope, z_out, s);
}

int pick_t(int64_t n) { return n > 64 ? 8 : (n > 32 ? 4 : 2); }

struct DwPlan {
  int nsplit;
  int64_t rows_per_split;
};
DwPlan dw_plan(int64_t R, int64_t M, int64_t N) {
  const int64_t tiles = ceil_div(M, 16 * pick_t(M)) * ceil_div(N, 16 * pick_t(N));
  int64_t want = (2 * kNumSMs + tiles - 1) / tiles;  // ~2 CTAs per SM in total
  int64_t chunks = ceil_div(R, BK);
  int64_t nsplit = want < 1 ? 1 : want;
  if (nsplit > chunks) nsplit = chunks < 1 ? 1 : chunks;
  int64_t rps = ceil_div(ceil_div(R, nsplit), BK) * BK;
  if (rps < BK) rps = BK;
  nsplit = ceil_div(R, rps);
  if (nsplit < 1) nsplit = 1;
  return {(int)nsplit, rps};
}

template <int TM, int TN>
void launch_tn(const float* A, const float* B, float* part, float* pcs, int64_t R, int M, int N, const DwPlan& pl,
               cudaStream_t s) {
  const bool av = (M % 4 == 0) && al16(A), bv = (N % 4 == 0) && al16(B);
  dim3 grid((unsigned)pl.nsplit, (unsigned)ceil_div(M, 16 * TM), (unsigned)ceil_div(N, 16 * TN));
  if (av && bv) gemm_tn_kernel<TM, TN, 4, 4><<<grid, kThreads, 0, s>>>(A, B, part, pcs, R, M, N, pl.rows_per_split);
  else if (av) gemm_tn_kernel<TM, TN, 4, 1><<<grid, kThreads, 0, s>>>(A, B, part, pcs, R, M, N, pl.rows_per_split);
  else if (bv) gemm_tn_kernel<TM, TN, 1, 4><<<grid, kThreads, 0, s>>>(A, B, part, pcs, R, M, N, pl.rows_per_split);
  else gemm_tn_kernel<TM, TN, 1, 1><<<grid, kThreads, 0, s>>>(A, B, part, pcs, R, M, N, pl.rows_per_split);
}

template <int TM>
void launch_tn_m(int tn, const float* A, const float* B, float* part, float* pcs, int64_t R, int M, int N,
                 const DwPlan& pl, cudaStream_t s) {
  if (tn == 8) launch_tn<TM, 8>(A, B, part, pcs, R, M, N, pl, s);
  else if (tn == 4) launch_tn<TM, 4>(A, B, part, pcs, R, M, N, pl, s);
  else launch_tn<TM, 2>(A, B, part, pcs, R, M, N, pl, s);
}

struct ColsumPlan {
  int nblk;
  int64_t rows_per_block;
};
ColsumPlan colsum_plan(int64_t R) {
  int64_t nblk = 4 * kNumSMs;
  int64_t rpb = ceil_div(R, nblk);
  if (rpb < 8) rpb = 8;
  nblk = ceil_div(R, rpb);
  if (nblk < 1) nblk = 1;
  return {(int)nblk, rpb};
}

}  // namespace
}  // namespace gcl

namespace gcl {
int umma_linear(const float* A, const float* W_nk, float* C, int64_t M, int64_t N, int64_t K, const float* bias,
                const float* slope, float* z_out, cudaStream_t s, const float* z_in = nullptr,
                const float* act_slope = nullptr, float* dslope_part = nullptr, int* n_parts = nullptr,
                const float* att = nullptr, float* sc_src = nullptr, float* sc_dst = nullptr,
                float* colsum_part = nullptr, int* n_colsum_parts = nullptr);
                                                                     // umma_gemm.cu
int umma_dw_splits(int64_t R, int64_t M, int64_t N);
int umma_dw(const float* A, const float* B, float* part, float* part_colsum, int64_t R, int64_t M, int64_t N,
            cudaStream_t s, int* n_bias_parts);
static int g_dense_mode = GCL_DENSE_AUTO;
constexpr int64_t kUmmaMinRows = 2048;   // below this the FFMA kernel's many small CTAs win
}  // namespace gcl

using namespace gcl;

extern "C" int gcl_linear_fwd_f32(const float* x, const float* W, const float* bias, float* y, int64_t rows,
                                  int64_t c_in, int64_t c_out, const float* prelu_slope, float* z_out,
                                  float* wt_scratch, void* stream) {
  GCL_CHECK_ARG(x && W && y && wt_scratch, "gcl_linear_fwd_f32: null pointer argument");
  GCL_CHECK_ARG(rows >= 0 && c_in > 0 && c_out > 0 && c_in <= 65536 && c_out <= 65536,
                "gcl_linear_fwd_f32: bad sizes rows=%lld c_in=%lld c_out=%lld", (long long)rows, (long long)c_in,
                (long long)c_out);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows == 0) return GCL_OK;
  if (g_dense_mode != GCL_DENSE_FFMA && rows >= kUmmaMinRows) {   // tcgen05 3xTF32: W [Cout, Cin] is already N x K, K-major
    const int rc = umma_linear(x, W, y, rows, c_out, c_in, bias, prelu_slope, z_out, s);
    if (rc != GCL_ERR_UNSUPPORTED) return rc;
  }
  dim3 tg((unsigned)ceil_div(c_in, 32), (unsigned)ceil_div(c_out, 32));
  transpose_kernel<<<tg, dim3(32, 8), 0, s>>>(W, wt_scratch, (int)c_out, (int)c_in);
  GCL_CHECK_LAUNCH("gcl_linear_fwd_f32(transpose)");
  return gemm_nn(x, wt_scratch, y, rows, c_out, c_in, bias, prelu_slope, z_out, s);
}

extern "C" int gcl_gat_scores_f32(const float* z, const float* att_src, const float* att_dst, float* a_src, float* a_dst,
                                  int64_t rows, int64_t heads, int64_t c, void* stream);

// y = x W^T (the bias-free `lin` of a single-head GATConv) together with the node terms of the attention logits,
// a_src[r] = <y[r], att_src>, a_dst[r] = <y[r], att_dst>: one kernel on the tcgen05 path (GEMM epilogue), otherwise
// the GEMM followed by gcl_gat_scores_f32.
extern "C" int gcl_linear_fwd_scores_f32(const float* x, const float* W, float* y, const float* att_src,
                                         const float* att_dst, float* a_src, float* a_dst, int64_t rows, int64_t c_in,
                                         int64_t c_out, float* wt_scratch /* c_in*c_out + 2*c_out floats */,
                                         void* stream) {
  GCL_CHECK_ARG(x && W && y && att_src && att_dst && a_src && a_dst && wt_scratch,
                "gcl_linear_fwd_scores_f32: null pointer argument");
  GCL_CHECK_ARG(rows >= 0 && c_in > 0 && c_out > 0 && c_in <= 65536 && c_out <= 65536,
                "gcl_linear_fwd_scores_f32: bad sizes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows == 0) return GCL_OK;
  if (g_dense_mode != GCL_DENSE_FFMA && rows >= kUmmaMinRows) {
    // the kernel wants [att_src ; att_dst] contiguous: pack them behind the transpose scratch
    float* att = wt_scratch + c_in * c_out;
    cudaMemcpyAsync(att, att_src, sizeof(float) * c_out, cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(att + c_out, att_dst, sizeof(float) * c_out, cudaMemcpyDeviceToDevice, s);
    const int rc = umma_linear(x, W, y, rows, c_out, c_in, nullptr, nullptr, nullptr, s, nullptr, nullptr, nullptr,
                               nullptr, att, a_src, a_dst);
    if (rc != GCL_ERR_UNSUPPORTED) return rc;
  }
  const int rc = gcl_linear_fwd_f32(x, W, nullptr, y, rows, c_in, c_out, nullptr, nullptr, wt_scratch, stream);
  if (rc != GCL_OK) return rc;
  return gcl_gat_scores_f32(y, att_src, att_dst, a_src, a_dst, rows, 1, c_out, stream);
}

extern "C" int gcl_linear_bwd_dx_f32(const float* dy, const float* W, float* dx, int64_t rows, int64_t c_in,
                                     int64_t c_out, float* wt_scratch, void* stream) {
  GCL_CHECK_ARG(dy && W && dx && wt_scratch, "gcl_linear_bwd_dx_f32: null pointer argument");
  GCL_CHECK_ARG(rows >= 0 && c_in > 0 && c_out > 0 && c_in <= 65536 && c_out <= 65536,
                "gcl_linear_bwd_dx_f32: bad sizes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows == 0) return GCL_OK;
  if (g_dense_mode != GCL_DENSE_FFMA && rows >= kUmmaMinRows) {   // tcgen05 wants N x K, K-major: W^T [Cin, Cout]
    dim3 tg((unsigned)ceil_div(c_in, 32), (unsigned)ceil_div(c_out, 32));
    transpose_kernel<<<tg, dim3(32, 8), 0, s>>>(W, wt_scratch, (int)c_out, (int)c_in);
    GCL_CHECK_LAUNCH("gcl_linear_bwd_dx_f32(transpose)");
    const int rc = umma_linear(dy, wt_scratch, dx, rows, c_in, c_out, nullptr, nullptr, nullptr, s);
    if (rc != GCL_ERR_UNSUPPORTED) return rc;
  }
  return gemm_nn(dy, W, dx, rows, c_in, c_out, nullptr, nullptr, nullptr, s);
}

extern "C" size_t gcl_prelu_bwd_workspace_bytes(int64_t n);
extern "C" int gcl_prelu_bwd_f32(const float* dy, const float* x, const float* slope, float* dx, float* dslope, int64_t n,
                                 void* workspace, size_t workspace_bytes, void* stream);

extern "C" size_t gcl_linear_bwd_dx_prelu_workspace_bytes(int64_t rows, int64_t c_in) {
  if (rows < 0 || c_in <= 0) return 0;
  // slope partials (one per CTA) + column-sum partials (two per CTA) | or the two-kernel fallback's needs
  const size_t a = gcl_prelu_bwd_workspace_bytes(rows * c_in), c = gcl_colsum_workspace_bytes(rows, c_in);
  const size_t b = (size_t)kNumSMs * sizeof(float) * (1 + 2 * (size_t)c_in) + 256;
  return (a > c ? a : c) > b ? (a > c ? a : c) : b;
}

// dz_in = (dy W) * PReLU'(z_in), dslope = sum((dy W) * min(z_in, 0)): the backward of "PReLU then Linear" w.r.t. the
// PReLU's input, in one kernel on the tcgen05 path (epilogue of the dX GEMM), else dX followed by the in-place
// PReLU backward.
extern "C" int gcl_linear_bwd_dx_prelu_f32(const float* dy, const float* W, const float* z_in, const float* slope,
                                           float* dz_in, float* dslope, float* dcolsum, int64_t rows, int64_t c_in,
                                           int64_t c_out, float* wt_scratch, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  GCL_CHECK_ARG(dy && W && z_in && slope && dz_in && dslope && wt_scratch && workspace,
                "gcl_linear_bwd_dx_prelu_f32: null pointer argument");
  GCL_CHECK_ARG(rows >= 0 && c_in > 0 && c_out > 0 && c_in <= 65536 && c_out <= 65536,
                "gcl_linear_bwd_dx_prelu_f32: bad sizes");
  if (workspace_bytes < gcl_linear_bwd_dx_prelu_workspace_bytes(rows, c_in)) {
    set_error("gcl_linear_bwd_dx_prelu_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    cudaMemsetAsync(dslope, 0, sizeof(float), s);
    if (dcolsum) cudaMemsetAsync(dcolsum, 0, sizeof(float) * c_in, s);
    return GCL_OK;
  }
  if (g_dense_mode != GCL_DENSE_FFMA && rows >= kUmmaMinRows) {
    dim3 tg((unsigned)ceil_div(c_in, 32), (unsigned)ceil_div(c_out, 32));
    transpose_kernel<<<tg, dim3(32, 8), 0, s>>>(W, wt_scratch, (int)c_out, (int)c_in);
    GCL_CHECK_LAUNCH("gcl_linear_bwd_dx_prelu_f32(transpose)");
    int n_parts = 0, n_cs = 0;
    float* part = static_cast<float*>(workspace);
    float* cs_part = part + kNumSMs;
    const int rc = umma_linear(dy, wt_scratch, dz_in, rows, c_in, c_out, nullptr, nullptr, nullptr, s, z_in, slope, part,
                               &n_parts, nullptr, nullptr, nullptr, dcolsum ? cs_part : nullptr, &n_cs);
    if (rc == GCL_OK) {
      reduce_partials_kernel<<<1, dim3(32, 32), 0, s>>>(part, dslope, 1, n_parts);
      GCL_CHECK_LAUNCH("gcl_linear_bwd_dx_prelu_f32(reduce)");
      if (dcolsum && n_cs > 0) {             // the wide-layer kernel summed the columns in its epilogue
        reduce_partials_kernel<<<(unsigned)ceil_div(c_in, 32), dim3(32, 32), 0, s>>>(cs_part, dcolsum, c_in, n_cs);
        GCL_CHECK_LAUNCH("gcl_linear_bwd_dx_prelu_f32(reduce colsum)");
        return GCL_OK;
      }
      return dcolsum ? gcl_colsum_f32(dz_in, dcolsum, rows, c_in, workspace, workspace_bytes, stream) : GCL_OK;
    }
    if (rc != GCL_ERR_UNSUPPORTED) return rc;
  }
  int rc = gcl_linear_bwd_dx_f32(dy, W, dz_in, rows, c_in, c_out, wt_scratch, stream);
  if (rc != GCL_OK) return rc;
  rc = gcl_prelu_bwd_f32(dz_in, z_in, slope, dz_in, dslope, rows * c_in, workspace, workspace_bytes, stream);
  if (rc != GCL_OK || !dcolsum) return rc;
  return gcl_colsum_f32(dz_in, dcolsum, rows, c_in, workspace, workspace_bytes, stream);
}

extern "C" int gcl_set_dense_mode(int mode) {
  GCL_CHECK_ARG(mode == GCL_DENSE_AUTO || mode == GCL_DENSE_FFMA, "gcl_set_dense_mode: bad mode %d", mode);
  g_dense_mode = mode;
  return GCL_OK;
}
extern "C" int gcl_get_dense_mode(void) { return g_dense_mode; }


extern "C" size_t gcl_linear_bwd_dw_workspace_bytes(int64_t rows, int64_t c_in, int64_t c_out) {
  if (rows < 0 || c_in <= 0 || c_out <= 0) return 0;
  DwPlan pl = dw_plan(rows, c_out, c_in);
  int nsplit = pl.nsplit;
  const int us = umma_dw_splits(rows, c_out, c_in);
  if (us > nsplit) nsplit = us;
  return (size_t)nsplit * (size_t)(c_out * c_in + 16 * c_out) * sizeof(float) + 256;   // + bias partials (<= 16 per slice)
}

extern "C" int gcl_linear_bwd_dw_f32(const float* dy, const float* x, float* dW, float* dbias, int64_t rows,
                                     int64_t c_in, int64_t c_out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  GCL_CHECK_ARG(dy && x && dW && workspace, "gcl_linear_bwd_dw_f32: null pointer argument");
  GCL_CHECK_ARG(rows >= 0 && c_in > 0 && c_out > 0 && c_in <= 65536 && c_out <= 65536,
                "gcl_linear_bwd_dw_f32: bad sizes");
  if (workspace_bytes < gcl_linear_bwd_dw_workspace_bytes(rows, c_in, c_out)) {
    set_error("gcl_linear_bwd_dw_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int M = (int)c_out, N = (int)c_in;
  if (rows == 0) {
    cudaMemsetAsync(dW, 0, sizeof(float) * M * N, s);
    if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * M, s);
    return GCL_OK;
  }
  float* part = static_cast<float*>(workspace);
  if (g_dense_mode != GCL_DENSE_FFMA && rows >= kUmmaMinRows) {   // tcgen05 3xTF32, rows split over <= 148 CTAs
    const int us = umma_dw_splits(rows, M, N);
    if (us > 0) {
      float* upcs = part + (size_t)us * M * N;
      int nbp = us;
      const int rc = umma_dw(dy, x, part, dbias ? upcs : nullptr, rows, M, N, s, &nbp);
      if (rc != GCL_OK) return rc;
      reduce_partials_kernel<<<(unsigned)ceil_div((int64_t)M * N, 32), dim3(32, 32), 0, s>>>(part, dW, (int64_t)M * N, us);
      GCL_CHECK_LAUNCH("gcl_linear_bwd_dw_f32(reduce)");
      if (dbias) {
        reduce_partials_kernel<<<(unsigned)ceil_div(M, 32), dim3(32, 32), 0, s>>>(upcs, dbias, M, nbp);
        GCL_CHECK_LAUNCH("gcl_linear_bwd_dw_f32(reduce bias)");
      }
      return GCL_OK;
    }
  }
  DwPlan pl = dw_plan(rows, M, N);
  float* pcs = part + (size_t)pl.nsplit * M * N;
  const int tm = pick_t(M), tn = pick_t(N);
  if (tm == 8) launch_tn_m<8>(tn, dy, x, part, dbias ? pcs : nullptr, rows, M, N, pl, s);
  else if (tm == 4) launch_tn_m<4>(tn, dy, x, part, dbias ? pcs : nullptr, rows, M, N, pl, s);
  else launch_tn_m<2>(tn, dy, x, part, dbias ? pcs : nullptr, rows, M, N, pl, s);
  GCL_CHECK_LAUNCH("gcl_linear_bwd_dw_f32(gemm_tn)");
  reduce_partials_kernel<<<(unsigned)ceil_div((int64_t)M * N, 32), dim3(32, 32), 0, s>>>(part, dW, (int64_t)M * N, pl.nsplit);
  GCL_CHECK_LAUNCH("gcl_linear_bwd_dw_f32(reduce)");
  if (dbias) {
    reduce_partials_kernel<<<(unsigned)ceil_div(M, 32), dim3(32, 32), 0, s>>>(pcs, dbias, M, pl.nsplit);
    GCL_CHECK_LAUNCH("gcl_linear_bwd_dw_f32(reduce bias)");
  }
  return GCL_OK;
}

extern "C" size_t gcl_colsum_workspace_bytes(int64_t rows, int64_t cols) {
  if (rows < 0 || cols <= 0) return 0;
  return (size_t)colsum_plan(rows).nblk * (size_t)cols * sizeof(float) + 256;
}

extern "C" int gcl_colsum_f32(const float* x, float* out, int64_t rows, int64_t cols, void* workspace,
                              size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(x && out && workspace, "gcl_colsum_f32: null pointer argument");
  GCL_CHECK_ARG(rows >= 0 && cols > 0 && cols <= (1 << 20), "gcl_colsum_f32: bad sizes");
  if (workspace_bytes < gcl_colsum_workspace_bytes(rows, cols)) {
    set_error("gcl_colsum_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    cudaMemsetAsync(out, 0, sizeof(float) * cols, s);
    return GCL_OK;
  }
  ColsumPlan pl = colsum_plan(rows);
  float* part = static_cast<float*>(workspace);
  if ((cols & 3) == 0 && cols <= 1024 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0)
    colsum_partial_v4_kernel<<<pl.nblk, 256, 0, s>>>(reinterpret_cast<const float4*>(x), part, rows, (int)cols,
                                                     pl.rows_per_block);
  else
    colsum_partial_kernel<<<pl.nblk, dim3(32, 8), 0, s>>>(x, part, rows, (int)cols, pl.rows_per_block);
  GCL_CHECK_LAUNCH("gcl_colsum_f32(partial)");
  reduce_partials_kernel<<<(unsigned)ceil_div(cols, 32), dim3(32, 32), 0, s>>>(part, out, cols, pl.nblk);
  GCL_CHECK_LAUNCH("gcl_colsum_f32(reduce)");
  return GCL_OK;
}

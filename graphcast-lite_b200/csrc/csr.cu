// K1: PyG edge_index -> receiver-grouped CSR + sender-grouped CSR, built on the device, stable
// (entries of a row are in ascending PyG edge position), deterministic, no host synchronisation.
//
// Stands in for what torch_geometric recomputes on every conv call:
//   GCNConv : add_remaining_self_loops + degree scatter + d^-1/2[src] w d^-1/2[dst]   (gcn_norm)
//   GATConv : remove_self_loops + add_self_loops
//   SimpleConv(mean): count per receiver
// Reference call sites: /root/reference/src/models.py:414,419,425,431.
#include "common.cuh"
#include "scan.cuh"

namespace gcl {
namespace {

__global__ void keep_flags_kernel(const int64_t* __restrict__ ei, int64_t E, int mode,
                                  int32_t* __restrict__ flag) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  flag[e] = (mode == GCL_CSR_LOOPS) ? (ei[e] != ei[E + e]) : 1;
}

// Writes the PyG-order edge list (kept edges, then loops) and raw weights; counts rows.
__global__ void emit_pyg_kernel(const int64_t* __restrict__ ei, const float* __restrict__ ew, int64_t E,
                                int64_t N, int mode, const int32_t* __restrict__ pos, int64_t cap,
                                int64_t* __restrict__ ei_out, float* __restrict__ w_pyg,
                                int32_t* __restrict__ cnt, int32_t* __restrict__ cnt_t,
                                int32_t* __restrict__ nnz_out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int32_t K = pos[E];
  if (i == 0) *nnz_out = K + (mode == GCL_CSR_LOOPS ? (int32_t)N : 0);
  if (i < E) {
    const int64_t s = ei[i], d = ei[E + i];
    const bool keep = (mode != GCL_CSR_LOOPS) || (s != d);
    if (keep) {
      const int32_t p = pos[i];
      ei_out[p] = s;
      ei_out[cap + p] = d;
      if (w_pyg) w_pyg[p] = ew ? ew[i] : 1.f;
      atomicAdd(&cnt[d], 1);
      atomicAdd(&cnt_t[s], 1);
    }
  } else if (mode == GCL_CSR_LOOPS && i < E + N) {
    const int64_t v = i - E;
    ei_out[K + v] = v;
    ei_out[cap + K + v] = v;
    if (w_pyg) w_pyg[K + v] = 1.f;
    atomicAdd(&cnt[v], 1);
    atomicAdd(&cnt_t[v], 1);
  }
}

// add_remaining_self_loops: an existing i->i edge hands its weight to the appended loop.
__global__ void loop_weight_kernel(const int64_t* __restrict__ ei, const float* __restrict__ ew, int64_t E,
                                   const int32_t* __restrict__ pos, float* __restrict__ w_pyg) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t s = ei[e];
  if (s == ei[E + e]) w_pyg[pos[E] + s] = ew[e];
}

__global__ void fill_kernel(const int64_t* __restrict__ ei_out, int64_t cap,
                            const int32_t* __restrict__ nnz_p, const int32_t* __restrict__ rowptr,
                            const int32_t* __restrict__ rowptr_t, int32_t* __restrict__ cursor,
                            int32_t* __restrict__ cursor_t, int32_t* __restrict__ tmp,
                            int32_t* __restrict__ tmp_t) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= *nnz_p) return;
  const int64_t s = ei_out[p], d = ei_out[cap + p];
  tmp[rowptr[d] + atomicAdd(&cursor[d], 1)] = (int32_t)p;
  tmp_t[rowptr_t[s] + atomicAdd(&cursor_t[s], 1)] = (int32_t)p;
}

// One warp per row: order the row's entries by PyG position (keys are unique) by ranking.
__global__ void row_sort_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ tmp,
                                int32_t* __restrict__ perm, int64_t N) {
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int32_t base = rowptr[row], len = rowptr[row + 1] - base;
  if (len <= 32) {
    const int32_t key = lane < len ? tmp[base + lane] : INT32_MAX;
    int rank = 0;
    for (int j = 0; j < len; ++j) rank += (__shfl_sync(0xffffffffu, key, j) < key);
    if (lane < len) perm[base + rank] = key;
  } else {
    for (int32_t i = lane; i < len; i += 32) {
      const int32_t key = tmp[base + i];
      int rank = 0;
      for (int32_t j = 0; j < len; ++j) rank += (tmp[base + j] < key);
      perm[base + rank] = key;
    }
  }
}

__global__ void gather_cols_kernel(const int64_t* __restrict__ ei_out, int64_t cap,
                                   const int32_t* __restrict__ nnz_p, const int32_t* __restrict__ perm,
                                   const int32_t* __restrict__ perm_t, int32_t* __restrict__ col,
                                   int32_t* __restrict__ col_t, int32_t* __restrict__ inv_perm) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= *nnz_p) return;
  const int32_t p = perm[k];
  col[k] = (int32_t)ei_out[p];
  inv_perm[p] = (int32_t)k;
  col_t[k] = (int32_t)ei_out[cap + perm_t[k]];
}

__global__ void t2r_kernel(const int32_t* __restrict__ nnz_p, const int32_t* __restrict__ perm_t,
                           const int32_t* __restrict__ inv_perm, int32_t* __restrict__ t2r) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= *nnz_p) return;
  t2r[k] = inv_perm[perm_t[k]];
}

// ---- weights -------------------------------------------------------------------------------------
__global__ void degree_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                              const float* __restrict__ w_pyg, int64_t N, float* __restrict__ dis) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  float deg = 0.f;
  for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) deg += w_pyg ? w_pyg[perm[k]] : 1.f;
  // torch: deg.pow(-0.5) then inf -> 0   (IEEE sqrt + divide, like the CPU path)
  const float r = __fdiv_rn(1.f, __fsqrt_rn(deg));
  dis[i] = isinf(r) ? 0.f : r;
}

__global__ void weights_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                               const int32_t* __restrict__ perm, const float* __restrict__ w_pyg,
                               const float* __restrict__ dis, int64_t N, int kind,
                               float* __restrict__ w_csr) {
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const int32_t b = rowptr[row], e = rowptr[row + 1];
  const float cnt = (float)max(e - b, 1);
  for (int32_t k = b + lane; k < e; k += 32) {
    const float w = w_pyg ? w_pyg[perm[k]] : 1.f;
    float v;
    if (kind == GCL_NORM_GCN) v = __fmul_rn(__fmul_rn(dis[col[k]], w), dis[row]);
    else if (kind == GCL_NORM_MEAN) v = __fdiv_rn(w, cnt);
    else v = w;
    w_csr[k] = v;
  }
}

__global__ void permute_weights_kernel(const int32_t* __restrict__ rowptr, int64_t N,
                                       const int32_t* __restrict__ t2r, const float* __restrict__ w_csr,
                                       float* __restrict__ w_csr_t) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= rowptr[N]) return;
  w_csr_t[k] = w_csr[t2r[k]];
}

struct CsrWs {
  int32_t *pos, *cnt, *cnt_t, *cursor, *cursor_t, *tmp, *tmp_t, *inv_perm, *flag;
  size_t bytes;
};

CsrWs carve(void* base, int64_t E, int64_t N, int64_t cap) {
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  char* p = static_cast<char*>(base);
  size_t off = 0;
  CsrWs w{};
  auto take = [&](size_t n_i32) {
    int32_t* r = reinterpret_cast<int32_t*>(p + off);
    off += up(n_i32 * sizeof(int32_t));
    return r;
  };
  // the four counters are contiguous so one memset clears them
  w.cnt = take(N + 1);
  w.cnt_t = take(N + 1);
  w.cursor = take(N + 1);
  w.cursor_t = take(N + 1);
  w.flag = take(E + 1);
  w.pos = take(E + 1);
  w.tmp = take(cap + 1);
  w.tmp_t = take(cap + 1);
  w.inv_perm = take(cap + 1);
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace gcl

using namespace gcl;

extern "C" size_t gcl_csr_workspace_bytes(int64_t E, int64_t N) {
  if (E < 0 || N < 0) return 0;
  return carve(nullptr, E, N, E + N).bytes;
}

extern "C" int gcl_csr_build(const int64_t* edge_index, const float* edge_weight, int64_t E, int64_t N,
                             int mode, int64_t* ei_out, float* w_pyg, int32_t* rowptr, int32_t* col,
                             int32_t* perm, int32_t* rowptr_t, int32_t* col_t, int32_t* perm_t,
                             int32_t* t2r, int32_t* nnz_out, void* workspace, size_t workspace_bytes,
                             void* stream) {
  GCL_CHECK_ARG(E >= 0 && N > 0, "gcl_csr_build: need num_edges >= 0 and num_nodes > 0 (got %lld, %lld)",
                (long long)E, (long long)N);
  GCL_CHECK_ARG(mode == GCL_CSR_RAW || mode == GCL_CSR_LOOPS, "gcl_csr_build: bad mode %d", mode);
  GCL_CHECK_ARG(E + N < (int64_t)INT32_MAX, "gcl_csr_build: graph too large for int32 indices");
  GCL_CHECK_ARG((E == 0 || edge_index) && ei_out && rowptr && col && perm && rowptr_t && col_t && perm_t &&
                    t2r && nnz_out && workspace,
                "gcl_csr_build: null pointer argument");
  GCL_CHECK_ARG(!(edge_weight && !w_pyg), "gcl_csr_build: edge_weight given but w_pyg is null");
  const int64_t cap = (mode == GCL_CSR_LOOPS) ? E + N : E;
  CsrWs w = carve(workspace, E, N, E + N);
  if (workspace_bytes < w.bytes) {
    set_error("gcl_csr_build: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int T = 256;
  cudaError_t e = cudaMemsetAsync(w.cnt, 0, (char*)w.flag - (char*)w.cnt, s);
  if (e != cudaSuccess) return fail_cuda(e, "gcl_csr_build(memset)");
  if (E > 0) {
    keep_flags_kernel<<<(unsigned)ceil_div(E, T), T, 0, s>>>(edge_index, E, mode, w.flag);
    GCL_CHECK_LAUNCH("gcl_csr_build(keep_flags)");
  }
  scan_exclusive_kernel<<<1, kScanThreads, 0, s>>>(w.flag, w.pos, E, nullptr);
  GCL_CHECK_LAUNCH("gcl_csr_build(scan pos)");
  emit_pyg_kernel<<<(unsigned)ceil_div(E + N, T), T, 0, s>>>(edge_index, edge_weight, E, N, mode, w.pos, cap,
                                                              ei_out, w_pyg, w.cnt, w.cnt_t, nnz_out);
  GCL_CHECK_LAUNCH("gcl_csr_build(emit)");
  if (edge_weight && mode == GCL_CSR_LOOPS && E > 0) {
    loop_weight_kernel<<<(unsigned)ceil_div(E, T), T, 0, s>>>(edge_index, edge_weight, E, w.pos, w_pyg);
    GCL_CHECK_LAUNCH("gcl_csr_build(loop weights)");
  }
  scan_exclusive_kernel<<<1, kScanThreads, 0, s>>>(w.cnt, rowptr, N, nullptr);
  GCL_CHECK_LAUNCH("gcl_csr_build(scan rows)");
  scan_exclusive_kernel<<<1, kScanThreads, 0, s>>>(w.cnt_t, rowptr_t, N, nullptr);
  GCL_CHECK_LAUNCH("gcl_csr_build(scan rows_t)");
  if (cap > 0) {
    fill_kernel<<<(unsigned)ceil_div(cap, T), T, 0, s>>>(ei_out, cap, nnz_out, rowptr, rowptr_t, w.cursor,
                                                         w.cursor_t, w.tmp, w.tmp_t);
    GCL_CHECK_LAUNCH("gcl_csr_build(fill)");
    row_sort_kernel<<<(unsigned)ceil_div(N * 32, T), T, 0, s>>>(rowptr, w.tmp, perm, N);
    GCL_CHECK_LAUNCH("gcl_csr_build(sort rows)");
    row_sort_kernel<<<(unsigned)ceil_div(N * 32, T), T, 0, s>>>(rowptr_t, w.tmp_t, perm_t, N);
    GCL_CHECK_LAUNCH("gcl_csr_build(sort rows_t)");
    gather_cols_kernel<<<(unsigned)ceil_div(cap, T), T, 0, s>>>(ei_out, cap, nnz_out, perm, perm_t, col, col_t,
                                                                w.inv_perm);
    GCL_CHECK_LAUNCH("gcl_csr_build(gather)");
    t2r_kernel<<<(unsigned)ceil_div(cap, T), T, 0, s>>>(nnz_out, perm_t, w.inv_perm, t2r);
    GCL_CHECK_LAUNCH("gcl_csr_build(t2r)");
  }
  return GCL_OK;
}

extern "C" int gcl_csr_weights(const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                               const int32_t* t2r, const float* w_pyg, int64_t N, int64_t nnz_cap,
                               int kind, float* deg_inv_sqrt, float* w_csr, float* w_csr_t,
                               void* stream) {
  GCL_CHECK_ARG(rowptr && col && perm && w_csr && N > 0, "gcl_csr_weights: null pointer / bad N");
  GCL_CHECK_ARG(kind >= GCL_NORM_NONE && kind <= GCL_NORM_MEAN, "gcl_csr_weights: bad kind %d", kind);
  GCL_CHECK_ARG(kind != GCL_NORM_GCN || deg_inv_sqrt, "gcl_csr_weights: GCN norm needs deg_inv_sqrt");
  GCL_CHECK_ARG(!w_csr_t || t2r, "gcl_csr_weights: w_csr_t needs t2r");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int T = 256;
  if (kind == GCL_NORM_GCN) {
    degree_kernel<<<(unsigned)ceil_div(N, T), T, 0, s>>>(rowptr, perm, w_pyg, N, deg_inv_sqrt);
    GCL_CHECK_LAUNCH("gcl_csr_weights(degree)");
  }
  weights_kernel<<<(unsigned)ceil_div(N * 32, T), T, 0, s>>>(rowptr, col, perm, w_pyg, deg_inv_sqrt, N, kind,
                                                            w_csr);
  GCL_CHECK_LAUNCH("gcl_csr_weights(weights)");
  if (w_csr_t && nnz_cap > 0) {
    permute_weights_kernel<<<(unsigned)ceil_div(nnz_cap, T), T, 0, s>>>(rowptr, N, t2r, w_csr, w_csr_t);
    GCL_CHECK_LAUNCH("gcl_csr_weights(permute)");
  }
  return GCL_OK;
}

/* gcl_b200 -- C ABI of the B200-native (sm_100a) message-passing kernels that replace the
 * torch_geometric arithmetic behind graphcast-lite's encode-process-decode hot path.
 *
 * The reference (ArturKKK/graphcast-lite) is pure Python and binds this path by IMPORT:
 *   /root/reference/src/models.py:21   from torch_geometric.nn import GCNConv, SimpleConv, GATConv, LayerNorm
 *   /root/reference/src/models.py:24   from torch_geometric.utils import dense_to_sparse, softmax
 *   /root/reference/src/models.py:220  from torch_geometric.utils import scatter
 * so the reference-side binding is the ctypes stub shown in INTEGRATION.md; every entry point below
 * names the reference/PyG call it stands in for.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller (PyTorch),
 *     never retained after return; the library allocates nothing on the device;
 *   - all functions only ENQUEUE on `stream` (a cudaStream_t passed as void*), never synchronise,
 *     and are CUDA-graph capturable;
 *   - return 0 on success, <0 on error (GCL_ERR_*); gcl_last_error() gives the thread-local text;
 *   - node-feature tensors are row-major fp32 [B, N, C] (B samples over one shared graph; B = 1 is the
 *     reference's layout [N, C]); `*_bstride` is the element distance between samples;
 *   - index arrays produced by gcl_csr_build are int32; edge_index is PyG's int64 [2, E]
 *     (row 0 = sender, row 1 = receiver).
 *   - results are deterministic: no floating-point atomics anywhere.
 */
#ifndef GCL_B200_H_
#define GCL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCL_OK 0
#define GCL_ERR_BAD_ARG (-1)
#define GCL_ERR_UNSUPPORTED (-2)
#define GCL_ERR_CUDA (-3)
#define GCL_ERR_WORKSPACE (-4)

/* ABI version (bumped on any signature change). */
int gcl_version(void);
/* Thread-local text of the last error returned on this host thread ("" if none). */
const char* gcl_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
long long gcl_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * K1  graph -> CSR.  Replaces, per call of GCNConv/GATConv/SimpleConv.forward, PyG's
 *     add_remaining_self_loops / remove_self_loops + add_self_loops / gcn_norm
 *     (torch_geometric/nn/conv/gcn_conv.py:gcn_norm; used via models.py:419,425,414) -- done ONCE per
 *     edge_index here and cached by the caller.
 *
 * mode: GCL_CSR_RAW   keep the edge list as it is                (SimpleConv, models.py:309)
 *       GCL_CSR_LOOPS drop i->i edges, append one i->i per node  (GCNConv / GATConv)
 * The "PyG edge order" is: kept edges in input order, then (LOOPS) nodes 0..N-1.  nnz = kept (+ N).
 *
 * Outputs (all device):
 *   ei_out   int64 [2, nnz_cap]  PyG-order edge list incl. loops (what GATConv returns with
 *                                return_attention_weights=True); nnz_cap = E + N (LOOPS) or E (RAW)
 *   w_pyg    fp32  [nnz_cap]     PyG-order raw edge weight (edge_weight or 1; loop = existing loop's
 *                                weight or 1)                      -- nullable
 *   rowptr   int32 [N+1], col int32 [nnz_cap], perm int32 [nnz_cap]
 *            receiver-grouped CSR: entries of row i = senders of i, ascending PyG position;
 *            perm[k] = PyG position of CSR entry k
 *   rowptr_t, col_t, perm_t      the same grouped by SENDER (col_t = receivers), for backward
 *   t2r      int32 [nnz_cap]     CSR position of sender-grouped entry k (t2r[k] = inv_perm[perm_t[k]])
 *   nnz_out  int32 [1]
 * edge_weight: nullable fp32 [E].
 */
#define GCL_CSR_RAW 0
#define GCL_CSR_LOOPS 1
size_t gcl_csr_workspace_bytes(int64_t num_edges, int64_t num_nodes);
int gcl_csr_build(const int64_t* edge_index, const float* edge_weight, int64_t num_edges,
                  int64_t num_nodes, int mode, int64_t* ei_out, float* w_pyg, int32_t* rowptr,
                  int32_t* col, int32_t* perm, int32_t* rowptr_t, int32_t* col_t, int32_t* perm_t,
                  int32_t* t2r, int32_t* nnz_out, void* workspace, size_t workspace_bytes,
                  void* stream);

/* Per-entry aggregation weights in CSR order (and the same values in sender-grouped order):
 *   GCL_NORM_GCN   w = d^-1/2[src] * w_pyg * d^-1/2[dst],  d[i] = sum of w_pyg over row i (inf -> 0)
 *                  (gcn_norm, improved=False)
 *   GCL_NORM_MEAN  w = w_pyg / max(count[dst], 1)           (SimpleConv(aggr="mean"), models.py:309)
 *   GCL_NORM_NONE  w = w_pyg
 * w_pyg nullable (= all ones).  Deterministic (row sums are sequential).
 */
#define GCL_NORM_NONE 0
#define GCL_NORM_GCN 1
#define GCL_NORM_MEAN 2
int gcl_csr_weights(const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                    const int32_t* t2r, const float* w_pyg, int64_t num_nodes, int64_t nnz_cap,
                    int kind, float* deg_inv_sqrt /* nullable [N] */, float* w_csr,
                    float* w_csr_t /* nullable */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2/K3/K6  deterministic segmented SpMM:  out[b,i,:] = epi( sum_{k in row i} w[k] * x[b,col[k],:] )
 *   forward of GCNConv.propagate / SimpleConv (index_select + mul + scatter_add_ in PyG) with the
 *   receiver-grouped CSR; their backward with the sender-grouped CSR (no atomics).
 *   epi: (+ bias[C]) then optional PReLU (slope = *prelu_slope, device scalar; models.py:316 shares
 *   one nn.PReLU per GraphLayer).  z_out (nullable) receives the value before PReLU.
 *   w nullable (= 1).  x and out must not alias.
 *   n_rows_out may be a prefix of the CSR's rows (the decoder only needs its grid rows, models.py:852);
 *   n_rows_in = rows x holds per sample: entries whose column is >= n_rows_in count as zero rows (the
 *   backward of such a prefix-restricted call reads a [B, n_rows_in, C] gradient through the transposed CSR).
 */
int gcl_spmm_f32(const int32_t* rowptr, const int32_t* col, const float* w, const float* x,
                 float* out, int64_t batch, int64_t n_rows_out, int64_t n_rows_in, int64_t channels, int64_t x_bstride,
                 int64_t out_bstride, const float* bias, const float* prelu_slope, float* z_out,
                 int64_t nnz /* number of CSR entries, 0 = unknown (scheduling hint only) */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2/K3/K6, tiled.  Same result as gcl_spmm_f32 (same per-row summation order), different schedule: the rows are
 *   grouped into TILES (<= max_rows rows whose union of source rows has <= max_union members); a CTA stages the
 *   union's feature rows in shared memory with asynchronous bulk copies (cp.async.bulk + mbarrier, no register
 *   staging) and reduces the tile's rows out of shared memory.  A feature row crosses L2 -> SM |union|/|rows|
 *   times instead of once per incident edge.  Rows with more distinct sources than max_union (the polar mesh
 *   rows of the 512x256 grid->mesh graph, grid_mesh_connectivity.py:53-104) are listed as HEAVY and reduced by a
 *   CTA-per-row kernel in a fixed order (no atomics).
 *
 * gcl_tile_plan_host builds the plan on the HOST from host copies of a CSR (rowptr/col of gcl_csr_build, either
 *   orientation).  order: nullable permutation of the rows -- tiles take rows in this order, so an order with
 *   spatial locality (the caller knows the mesh geometry) gives small unions; NULL = natural order.
 *   Only rows < n_rows_out get a tile; entries whose column is >= n_rows_in are masked (zero rows).
 *   pad_entries (1, 2 or 4): each row's entry list is padded to a multiple of it.
 *   Output arrays are sized by the caller for the worst case: tile_rowptr/tile_uptr [n_rows+1], rows [n_rows],
 *   eptr [n_rows+1], lidx/ek [nnz + (pad_entries-1)*n_rows], usrc [nnz], heavy_rows [n_rows], tile_desc [8*n_rows];
 *   counts[8] = {n_tiles, n_plan_rows, n_union_total, n_entries, n_heavy, max_rows, max_union, max_entries}.
 *   The caller uploads the used prefixes and fills a gcl_tile_plan with the DEVICE pointers.
 */
typedef struct gcl_tile_plan {
  const int32_t* tile_rowptr;  /* [n_tiles+1] first plan row of each tile                                */
  const int32_t* tile_uptr;    /* [n_tiles+1] first union entry of each tile                             */
  const int32_t* rows;         /* [n_plan_rows] CSR row handled by plan row i                            */
  const int32_t* eptr;         /* [n_plan_rows+1] first plan entry of plan row i                         */
  const uint16_t* lidx;        /* [n_entries] index into the tile's union, 0xFFFF = masked               */
  const int32_t* ek;           /* [n_entries] CSR position of the plan entry (weights / attention index) */
  const int32_t* usrc;         /* [n_union_total] source rows of each tile's union                       */
  const int32_t* heavy_rows;   /* [n_heavy] CSR rows outside the tiles (nullable if n_heavy == 0)        */
  const int32_t* tile_desc;    /* [n_tiles][8] {first plan row, rows, first union entry, union size, first plan
                                  entry, plan entries, 0, 0} -- what tile_rowptr/tile_uptr/eptr say, in one record */
  int32_t n_tiles, n_heavy, max_rows, max_union, max_entries;
  int32_t pad_entries;         /* every row's entry list is padded to a multiple of this (pad entries have
                                  lidx = 0xFFFF, ek = -1); the pipelined SpMM kernel wants 2, the attention kernels 1 */
} gcl_tile_plan;
int gcl_tile_plan_host(const int32_t* rowptr, const int32_t* col, const int32_t* order, int64_t n_rows,
                       int64_t n_rows_out, int64_t n_rows_in, int32_t max_rows, int32_t max_union,
                       int32_t max_entries, int32_t pad_entries, int32_t* tile_rowptr, int32_t* tile_uptr,
                       int32_t* rows, int32_t* eptr, uint16_t* lidx, int32_t* ek, int32_t* usrc,
                       int32_t* heavy_rows, int32_t* tile_desc, int64_t* counts);
/* plan: HOST pointer to a struct of DEVICE arrays, built with pad_entries = 2.  Persistent CTAs (one per SM) walk
 * the (tile, sample block) items through a multi-stage shared-memory ring: the rows and the tile's index data of item
 * i+2 are in flight (cp.async / cp.async.bulk) while item i is reduced.
 * ent: int32 [n_entries][2] = {lidx, weight bits} in plan order (the caller gathers w[ek] once per weight kind);
 *   pad and masked entries carry lidx = plan->max_union (the kernel's all-zero row), pads weight 0.  rowptr/col/w: the CSR the plan was built from, used for the heavy rows (w nullable = 1).
 * channels: multiple of 4, <= 128, rows 16-byte aligned (otherwise GCL_ERR_BAD_ARG: use gcl_spmm_f32). */
int gcl_spmm_tiled_f32(const gcl_tile_plan* plan, const int32_t* ent, const int32_t* rowptr, const int32_t* col,
                       const float* w, const float* x, float* out, int64_t batch, int64_t n_rows_in,
                       int64_t channels, int64_t x_bstride, int64_t out_bstride, const float* bias,
                       const float* prelu_slope, float* z_out, void* stream);

/* bf16 feature rows (north_star: "128-bit vectorised loads of the ... bf16/fp32 feature rows", tolerance rel 2e-2):
 * the same tiled engine with x / out / z_out holding bf16 ([B, N, C], strides in elements, C a multiple of 8, <= 256).
 * Half the bytes cross HBM and L2; weights, bias and accumulation stay fp32, results are rounded to nearest even.
 * Optional storage format of the aggregation layers, never the benchmark's headline dtype. */
int gcl_spmm_tiled_bf16(const gcl_tile_plan* plan, const int32_t* ent, const int32_t* rowptr, const int32_t* col,
                        const float* w, const void* x, void* out, int64_t batch, int64_t n_rows_in, int64_t channels,
                        int64_t x_bstride, int64_t out_bstride, const float* bias, const float* prelu_slope,
                        void* z_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7  node-wise dense transform  y = act(x W^T + b)  (torch.nn.Linear in MLP, models.py:74-98, and the
 *   bias-free `lin` inside GCNConv/GATConv).  x [R, Cin], W [Cout, Cin], y [R, Cout].  Default engine:
 *   tcgen05 tensor cores in 3xTF32 (hi/lo operand split, fp32 accumulation in TMEM, TMA in and out) -- fp32-level
 *   accuracy (rms error vs fp64 2.9e-7).  Layers with 64 < Cout <= 128 and Cin <= 128 run weights-stationary: W is
 *   the M-side operand held in tensor memory and the transposed tile W x^T is accumulated (dW likewise keeps dY^T
 *   in tensor memory); narrower layers keep W in shared memory.  rows < 2048 or widths that are not a multiple of
 *   4 floats, and gcl_set_dense_mode(GCL_DENSE_FFMA), take the fp32 FFMA kernels.
 *   bias / prelu_slope / z_out nullable (z_out = value before PReLU).
 *   wt_scratch: device scratch of Cin*Cout floats (holds the split / transposed W).
 */
int gcl_linear_fwd_f32(const float* x, const float* W, const float* bias, float* y, int64_t rows,
                       int64_t c_in, int64_t c_out, const float* prelu_slope, float* z_out,
                       float* wt_scratch, void* stream);
/* dx[R, Cin] = dy[R, Cout] W.  wt_scratch: Cin*Cout floats (holds W^T for the tensor-core path). */
/* `lin` of a single-head GATConv together with the node terms of the attention logits (PyG: (x * att).sum(-1)):
 * y = x W^T, a_src[r] = <y[r], att_src>, a_dst[r] = <y[r], att_dst>; wt_scratch: c_in*c_out + 2*c_out floats. */
int gcl_linear_fwd_scores_f32(const float* x, const float* W, float* y, const float* att_src,
                              const float* att_dst, float* a_src, float* a_dst, int64_t rows, int64_t c_in,
                              int64_t c_out, float* wt_scratch, void* stream);
int gcl_linear_bwd_dx_f32(const float* dy, const float* W, float* dx, int64_t rows, int64_t c_in,
                          int64_t c_out, float* wt_scratch, void* stream);
/* Backward of "PReLU, then Linear" with respect to the PReLU's input z_in [rows, c_in] (the MLP of models.py:74-98
 * alternates them): dz_in = (dy W) * PReLU'(z_in), *dslope = sum((dy W) * min(z_in, 0)).  One kernel on the
 * tcgen05 path (epilogue of the dX GEMM); otherwise dX followed by an in-place PReLU backward.
 * dcolsum (nullable, [c_in]): column sums of dz_in -- the bias gradient of a layer whose bias is added just before
 * the PReLU (GCNConv: models.py:419-424); the wide-layer tcgen05 kernel accumulates them in its epilogue. */
size_t gcl_linear_bwd_dx_prelu_workspace_bytes(int64_t rows, int64_t c_in);
int gcl_linear_bwd_dx_prelu_f32(const float* dy, const float* W, const float* z_in, const float* slope,
                                float* dz_in, float* dslope, float* dcolsum, int64_t rows, int64_t c_in,
                                int64_t c_out, float* wt_scratch, void* workspace, size_t workspace_bytes,
                                void* stream);
/* Which engine runs the dense transforms:
 *   GCL_DENSE_AUTO  tcgen05 tensor cores in 3xTF32 (hi/lo split, fp32 accumulate in TMEM; fp32-level
 *                   accuracy) when the shape fits (rows >= 2048, Cout <= 256), else FFMA   [default]
 *   GCL_DENSE_FFMA  CUDA-core fp32 FFMA kernels only (A/B comparison for the ncu evidence). */
#define GCL_DENSE_AUTO 0
#define GCL_DENSE_FFMA 1
int gcl_set_dense_mode(int mode);
int gcl_get_dense_mode(void);
/* dW[Cout, Cin] = dy^T x ; dbias[Cout] = column sums of dy (nullable).  Deterministic split over rows. */
size_t gcl_linear_bwd_dw_workspace_bytes(int64_t rows, int64_t c_in, int64_t c_out);
int gcl_linear_bwd_dw_f32(const float* dy, const float* x, float* dW, float* dbias, int64_t rows,
                          int64_t c_in, int64_t c_out, void* workspace, size_t workspace_bytes,
                          void* stream);

/* Column sums of a [R, C] matrix (bias gradients), deterministic.  workspace >= gcl_colsum_workspace_bytes. */
size_t gcl_colsum_workspace_bytes(int64_t rows, int64_t cols);
int gcl_colsum_f32(const float* x, float* out, int64_t rows, int64_t cols, void* workspace,
                   size_t workspace_bytes, void* stream);

/* PReLU (torch.nn.PReLU with one slope; models.py:78,159):  y = x > 0 ? x : a x */
int gcl_prelu_fwd_f32(const float* x, const float* slope, float* y, int64_t n, void* stream);
/* dx = dy * (x > 0 ? 1 : a);  dslope[1] = sum(dy * x * [x <= 0]).  Deterministic. */
size_t gcl_prelu_bwd_workspace_bytes(int64_t n);
int gcl_prelu_bwd_f32(const float* dy, const float* x, const float* slope, float* dx, float* dslope,
                      int64_t n, void* workspace, size_t workspace_bytes, void* stream);

/* The same, fused with the column sums of dx: dbias[C] = sum_rows dx (the bias gradient of the conv whose
 * output fed this PReLU; models.py:323-328 alternates GCNConv and the shared PReLU).  dy, x, dx [R, C]. */
size_t gcl_prelu_bwd_colsum_workspace_bytes(int64_t rows, int64_t c);
int gcl_prelu_bwd_colsum_f32(const float* dy, const float* x, const float* slope, float* dx,
                             float* dslope, float* dbias, int64_t rows, int64_t c, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a6  torch_geometric.nn.LayerNorm (models.py:103,370).
 *   mode "node" = F.layer_norm(x, (C,), gamma, beta, eps), the mode of every BASELINE config.
 *   (mode "graph" is composed on the host from these primitives; see gcl_b200/nn.)
 *   x, y [R, C]; mean/rstd [R] saved for backward; gamma/beta nullable together (affine=False).
 */
int gcl_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, float* y,
                          float* mean, float* rstd, int64_t rows, int64_t c, float eps,
                          void* stream);
size_t gcl_layernorm_bwd_workspace_bytes(int64_t rows, int64_t c);
int gcl_layernorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* mean,
                          const float* rstd, float* dx, float* dgamma, float* dbeta, int64_t rows,
                          int64_t c, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4/K5  GATConv (models.py:336-357) / SparseGATConv (models.py:112-151) after the dense transform:
 *   z [B, N, H, C] = lin(x);  a_s = <z, att_src>, a_d = <z, att_dst>   [B, N, H]
 *   e_ij = LeakyReLU_slope(a_s[j] + a_d[i]);  alpha = exp(e - max_i) / (sum_i exp(.) + 1e-16)
 *   out[b,i,:] = (concat ? [o_1..o_H] : mean_h o_h) + bias,  o_h = sum_j alpha_ijh z[b,j,h,:]
 * gcl_gat_fwd_f32: heads > 1 -- logits + LeakyReLU + segment softmax + aggregation + head mean/concat in ONE
 *   kernel; heads == 1 -- a coefficient kernel (thread per (sample, receiver)) followed by the SpMM kernel with
 *   per-sample weights (bias / PReLU in its epilogue): two launches, measured faster than the fused one
 *   (DESIGN.md 3).  gcl_gat_fwd_tiled_f32 (below) is the single-kernel heads == 1 forward on a tile plan.
 *   alpha_csr  fp32 [B, nnz, H]  attention in CSR order (kept for backward)
 *   alpha_pyg  nullable fp32 [B, nnz, H] attention in PyG edge order (return_attention_weights=True)
 */
int gcl_gat_scores_f32(const float* z, const float* att_src, const float* att_dst, float* a_src,
                       float* a_dst, int64_t rows /* B*N */, int64_t heads, int64_t c, void* stream);
int gcl_gat_fwd_f32(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const float* z,
                    const float* a_src, const float* a_dst, const float* bias, float* out,
                    float* alpha_csr, float* alpha_pyg,
                    const float* prelu_slope /* nullable: PReLU fused behind the bias (models.py:316, 426-428) */,
                    float* z_out /* nullable: the value before that PReLU; heads == 1 only */,
                    int64_t batch, int64_t n_nodes, int64_t nnz,
                    int64_t heads, int64_t c, int concat, float negative_slope, void* stream);
/* Backward.  Pass 1 (receiver-grouped): g = d(pre-LeakyReLU logit) per entry, da_dst.
 * Pass 2 (sender-grouped): dz = sum_i alpha_ij do_i + da_src att_src + da_dst att_dst, da_src.
 *   dout [B, N, Cout]; g_csr scratch [B, nnz, H]; da_src/da_dst [B, N, H] (outputs, also needed for
 *   d att_src = sum da_src z, computed by the caller with gcl_gat_datt_f32).
 */
int gcl_gat_bwd_f32(const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                    const int32_t* col_t, const int32_t* t2r, const float* z, const float* a_src,
                    const float* a_dst, const float* alpha_csr, const float* att_src,
                    const float* att_dst, const float* dout, float* g_csr, float* da_src,
                    float* da_dst, float* dz, int64_t batch, int64_t n_nodes, int64_t nnz,
                    int64_t heads, int64_t c, int concat, float negative_slope, void* stream);
/* heads == 1 GATConv on tile plans (see gcl_tile_plan): ONE forward kernel -- the tile CTA stages the union's z rows
 * by bulk copies, computes the attention coefficients of its rows (logits, LeakyReLU, segment softmax; written to
 * alpha_csr [B, nnz] and optionally alpha_pyg) while the copies are in flight, then aggregates out of shared
 * memory (+ bias, optional PReLU with z_out = pre-activation).  plan: receiver-grouped, no heavy rows. */
int gcl_gat_fwd_tiled_f32(const gcl_tile_plan* plan, const int32_t* perm, const float* z, const float* a_src,
                          const float* a_dst, const float* bias, float* out, float* alpha_csr, float* alpha_pyg,
                          const float* prelu_slope, float* z_out, int64_t batch, int64_t n_nodes, int64_t nnz,
                          int64_t c, float negative_slope, void* stream);
/* heads == 1 backward on tile plans: pass 1 on the receiver-grouped plan (z rows of the union and the tile's dout
 * rows in shared memory), pass 2 on the sender-grouped plan (dout rows of the union in shared memory).  Same
 * outputs and per-row summation order as gcl_gat_bwd_f32. */
int gcl_gat_bwd_tiled_f32(const gcl_tile_plan* plan, const gcl_tile_plan* plan_t, const int32_t* t2r, const float* z,
                          const float* a_src, const float* a_dst, const float* alpha_csr, const float* att_src,
                          const float* att_dst, const float* dout, float* g_csr, float* da_src, float* da_dst,
                          float* dz, int64_t batch, int64_t n_nodes, int64_t nnz, int64_t c, float negative_slope,
                          void* stream);
/* heads == 1 GATConv on the persistent warp-specialised engine (the one behind gcl_spmm_tiled_f32).  The attention
 * coefficients are kept in PLAN order, so a tile's coefficients are contiguous and travel with its rows:
 *   alpha_f [B, ef] receiver-grouped plan order; alr_f [B, ef] = alpha * LeakyReLU'(logit);
 *   alpha_t [B, et] sender-grouped plan order (ZERO-INITIALISED by the caller: plan pads are never written);
 *   g_t     [B, et] d(logit) in sender-grouped plan order (scratch, ZERO-INITIALISED by the caller);
 *   f2t int32 [ef]: sender-grouped plan entry of receiver-grouped plan entry e (-1 for pads);
 *   pcol int32 [ef]: sender of receiver-grouped plan entry e (col[ek[e]], -1 for pads);
 *   ent / ent_t: packed plan entries as for gcl_spmm_tiled_f32 (the weight field is ignored);
 *   plan / plan_t: pad_entries = 2, no heavy rows, every node in exactly one tile.
 * Forward = coefficient kernel (thread per (sample, row); writes alpha_f, alr_f, alpha_t and optionally alpha_pyg)
 *   + aggregation with per-sample weights (+ bias, optional PReLU with z_out = pre-activation).
 * Backward = pass 1 on the receiver-grouped plan (g_t, da_dst) + pass 2 on the sender-grouped plan (dz, da_src);
 *   same outputs as gcl_gat_bwd_f32.  gcl_gat_ws_supported: 1 if the tiles fit the shared-memory ring. */
int gcl_gat_ws_supported(const gcl_tile_plan* plan, const gcl_tile_plan* plan_t, int64_t c, int64_t batch);
int gcl_gat_fwd_ws_f32(const gcl_tile_plan* plan, const int32_t* ent, const int32_t* pcol, const int32_t* perm,
                       const int32_t* f2t, const float* z, const float* a_src, const float* a_dst, const float* bias,
                       float* out, float* alpha_f, float* alr_f, float* alpha_t, float* alpha_pyg,
                       const float* prelu_slope, float* z_out, int64_t batch, int64_t n_nodes, int64_t nnz, int64_t c,
                       int64_t ef, int64_t et, float negative_slope, void* stream);
int gcl_gat_bwd_ws_f32(const gcl_tile_plan* plan, const int32_t* ent, const gcl_tile_plan* plan_t, const int32_t* ent_t,
                       const int32_t* f2t, const float* z, const float* alpha_f, const float* alr_f,
                       const float* alpha_t, const float* att_src, const float* att_dst, const float* dout, float* g_t,
                       float* da_src, float* da_dst, float* dz, int64_t batch, int64_t n_nodes, int64_t c, int64_t ef,
                       int64_t et, void* stream);
/* datt_src[h,c] = sum_{b,n} da_src[b,n,h] z[b,n,h,c]  (same for dst).  Deterministic. */
size_t gcl_gat_datt_workspace_bytes(int64_t rows, int64_t heads, int64_t c);
int gcl_gat_datt_f32(const float* z, const float* da_src, const float* da_dst, float* datt_src,
                     float* datt_dst, int64_t rows, int64_t heads, int64_t c, void* workspace,
                     size_t workspace_bytes, void* stream);
/* SparseGATConv pruning (models.py:138-149): keep PyG-order edges with alpha >= threshold.
 * Writes the compacted int64 [2, kept] list to ei_kept (capacity nnz) and kept to count_out[1]. */
size_t gcl_edge_prune_workspace_bytes(int64_t nnz);
int gcl_edge_prune(const int64_t* ei_pyg, const float* alpha_pyg, int64_t nnz, int64_t ei_stride,
                   float threshold, int64_t* ei_kept, int64_t kept_stride, int32_t* count_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a10/a12  graph construction searches (one-time, fp64, fixed operation order, no FP contraction).
 *
 * Grid -> mesh radius query (create_graphs.py:126-153 -> grid_mesh_connectivity.py:53-104, scipy
 * cKDTree.query_ball_point): mesh vertex m is a hit of grid point g when
 * ((dx^2 + dy^2) + dz^2) <= radius^2 in fp64.  Two calls: _count fills offsets[G+1] (exclusive scan of
 * the per-point hit counts; offsets[G] = E), the caller allocates [2, E] and _fill writes
 * row 0 = g, row 1 = mesh_index_offset + m, senders ascending, receivers ascending within a sender
 * (the reference's receiver order inside a sender is cKDTree traversal order: compare as sets).
 * grid_xyz fp64 [G,3], mesh_xyz fp32 [M,3].
 */
size_t gcl_radius_query_workspace_bytes(int64_t n_grid);
int gcl_radius_query_count(const double* grid_xyz, const float* mesh_xyz, int64_t n_grid,
                           int64_t n_mesh, double radius, int32_t* offsets, void* workspace,
                           size_t workspace_bytes, void* stream);
int gcl_radius_query_fill(const double* grid_xyz, const float* mesh_xyz, int64_t n_grid,
                          int64_t n_mesh, double radius, const int32_t* offsets,
                          int64_t* edge_index_out, int64_t num_edges, int64_t mesh_index_offset,
                          void* stream);
/* Mesh -> grid containing triangle (create_graphs.py:271-289 -> grid_mesh_connectivity.py:139-184,
 * trimesh.proximity.closest_point): face_out[g] = id of the mesh triangle closest to grid point g
 * (Ericson closest point, tol.zero 1e-13; best two by squared distance, ties to the lower face id;
 * normal-alignment rule when both exceed tol.merge 1e-8 and differ by less).  Faces whose first
 * vertex is farther than prefilter_radius from the point are skipped (pass >= 3 max edge lengths).
 */
int gcl_closest_face(const double* grid_xyz, const float* mesh_xyz, const int32_t* faces,
                     int64_t n_grid, int64_t n_mesh, int64_t n_faces, double prefilter_radius,
                     int32_t* face_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a9/a14  glue of one forecast / training step (models.py:776-806, train.py:85-102,203-213; Adam).
 */
/* enc_in[b, n, :] = n < G ? [x[b,n,:TF], grid_static[n,:S], 0..] : [0..0, mesh_static[n-G,:S], 0..];
 * rows are out_width >= TF + S floats wide, the tail is zero (lets the caller keep rows 16-byte aligned). */
int gcl_assemble_input_f32(const float* x, const float* grid_static, const float* mesh_static,
                           float* enc_in, int64_t batch, int64_t n_grid, int64_t n_mesh, int64_t tf,
                           int64_t s, int64_t out_width, void* stream);
/* Node-axis concat / split of [B, N, C] tensors, one pass each (models.py:841-842 slices the encoder output into
 * grid / mesh rows, :865 concatenates grid rows and processed mesh rows; each is also the other's backward).
 *   concat: out[b, :Na] = a[b], out[b, Na:] = b_[b].   split: a[b] = x[b, :Na], b_[b] = x[b, Na:].
 *   a or b_ may be NULL in split (that part is not needed) and in concat (that part of out is zero-filled). */
int gcl_rows_concat_f32(const float* a, const float* b_, float* out, int64_t batch, int64_t na, int64_t nb,
                        int64_t c, void* stream);
int gcl_rows_split_f32(const float* x, float* a, float* b_, int64_t batch, int64_t na, int64_t nb, int64_t c,
                       void* stream);
/* One block of node rows of every sample: dst[b, dst_row0 + r, :] = src[b, src_row0 + r, :], r < rows, for
 * src [B, src_rows, C] and dst [B, dst_rows, C].  With it the host takes the mesh rows out of the encoder output
 * (models.py:842) and writes the processed mesh rows back IN PLACE (the torch.cat of models.py:865 without
 * moving the grid rows). */
int gcl_rows_block_copy_f32(const float* src, float* dst, int64_t batch, int64_t rows, int64_t c,
                            int64_t src_row0, int64_t src_rows, int64_t dst_row0, int64_t dst_rows, void* stream);
/* Residual + latitude-weighted MSE (train.py:85-102, 203-213), forward and gradient in one pass:
 *   out = (x_last ? x_last : 0) + delta;  loss = scale * inv_wsum * sum(w (out - y)^2),  w[b,g,c] = lat_w[g]
 *   (lat_w nullable = 1; inv_wsum = 1 / sum of all weights, host-computed);
 *   d_delta = scale * 2 w (out - y) * inv_wsum.   x_last / y rows have strides xl_stride / y_stride
 *   (views into [.., steps, C]).  loss_out[1] is overwritten, or added to when `accumulate`;
 *   out_state (nullable [B,G,C]) receives `out`; d_delta nullable. */
size_t gcl_wmse_workspace_bytes(int64_t batch, int64_t n_grid, int64_t c);
int gcl_wmse_f32(const float* delta, const float* x_last, int64_t xl_stride, const float* y,
                 int64_t y_stride, const float* lat_w, float inv_wsum, float* out_state,
                 float* d_delta, float* loss_out, int accumulate, float scale, int64_t batch,
                 int64_t n_grid, int64_t c, void* workspace, size_t workspace_bytes, void* stream);
/* y [rows, c_out] = x [rows, c_in] truncated or zero-padded along the channel axis: drops the alignment padding of
 * a layer computed 4-aligned (c_out < c_in) and, as its backward, re-pads the gradient (c_out > c_in). */
int gcl_resize_channels_f32(const float* x, float* y, int64_t rows, int64_t c_in, int64_t c_out, void* stream);
/* One autoregressive training step behind the model (train.py:201-227) in one pass over [B, G, C]:
 *   out = (residual ? state[b,g,obs-1,:] : 0) + delta            (train.py:203-207)
 *   loss += scale * sum w (out - y)^2 * inv_wsum,  w = node_w[g] * chan_w[c]   (weighted_mse_loss, train.py:85-102:
 *           node_w = latitude weights x spatial mask, chan_w = channel mask; inv_wsum = 1 / max(sum of all w, 1e-12))
 *   g_loss = d loss / d out  [B, G, C]
 *   new_state [B, G, obs, C] (nullable) = window slide [state[.., 1:, :], out'] with the carry-forward of
 *           train.py:218-227: out'[c] = state[.., obs-1, c] where carry[c] == 1 (static channel), y[c] where
 *           carry[c] == 2 (forcing channel), out[c] otherwise.  carry nullable (= all 0).
 *   y nullable (inference rollout without ground truth: no forcing values, the loss is meaningless).
 * gcl_ar_step_bwd_f32: given d loss (device scalar, nullable = 1) and d new_state (nullable), d_delta [B, G, C] and
 *   d_state [B, G, obs, C] (nullable).  Workspace as gcl_wmse_f32. */
int gcl_ar_step_f32(const float* delta, const float* state, const float* y, int64_t y_stride,
                    const float* node_w, const float* chan_w, const int32_t* carry, int residual,
                    float inv_wsum, float scale, float* new_state, float* g_loss, float* loss_out,
                    int accumulate, int64_t batch, int64_t n_grid, int64_t obs, int64_t c, void* workspace,
                    size_t workspace_bytes, void* stream);
int gcl_ar_step_bwd_f32(const float* g_loss, const float* dloss, const float* d_new_state, const int32_t* carry,
                        int residual, float* d_delta, float* d_state, int64_t batch, int64_t n_grid, int64_t obs,
                        int64_t c, void* stream);
/* ------------------------------------------------------------------------------------------------
 * f3  InteractionNet processor (models.py:166-285) and the v2 configs: what they need beyond the kernels above.
 *   activations of _get_activation (models.py:154-163) other than PReLU: ReLU, SiLU ("swish");
 *   torch_geometric LayerNorm(mode="graph") (edge_norm, models.py:201): per sample,
 *     y = (x - mean) / (std(unbiased=False) + eps) * gamma[c] + beta[c], statistics over ALL elems_per_sample
 *     elements (float64 accumulation, fixed order); stats float [batch][3] = (mean, 1 / (std + eps), std) is saved
 *     for the backward, which returns dx, coef (scratch float [batch][2]) and t_out = dy * xhat (nullable) whose
 *     column sums are d gamma (d beta = column sums of dy; both with gcl_colsum_f32);
 *   gcl_add_f32: the residual connections (models.py:226-227).
 */
#define GCL_ACT_RELU 1
#define GCL_ACT_SILU 2
int gcl_act_fwd_f32(const float* x, float* y, int64_t n, int kind, void* stream);
int gcl_act_bwd_f32(const float* dy, const float* x, float* dx, int64_t n, int kind, void* stream);
int gcl_add_f32(const float* a, const float* b, float* y, int64_t n, void* stream);
size_t gcl_layernorm_graph_workspace_bytes(int64_t batch);
int gcl_layernorm_graph_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* stats,
                                int64_t batch, int64_t elems_per_sample, int64_t c, float eps, void* workspace,
                                size_t workspace_bytes, void* stream);
int gcl_layernorm_graph_bwd_f32(const float* dy, const float* x, const float* gamma, const float* stats, float* dx,
                                float* t_out, float* coef, int64_t batch, int64_t elems_per_sample, int64_t c,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * f2  input pipeline, device side.  TimeseriesChunkDataset.__getitem__ (src/data/dataloader_chunked.py:179-223)
 *   converts, normalises and transposes every window on CPU workers; here the host only stages RAW windows and one
 *   kernel produces both model tensors for the whole batch:
 *   raw [B, W, n_lon, n_lat, f_total] (flat = 0) or [B, W, N, f_total] (flat = 1, pass n_lon = N, n_lat = 1) in the
 *   dataset's stored dtype; x_out [B, G, obs * f_used], y_out [B, G, (W - obs) * f_used], G lat-major
 *   (node = lat * n_lon + lon);  value = (float32(raw) - mean[f]) / std[f], the reference's float32 operations.
 */
#define GCL_RAW_F16 0
#define GCL_RAW_F32 1
#define GCL_RAW_F64 2
int gcl_window_assemble(const void* raw, int raw_dtype, const float* mean, const float* stdv, float* x_out,
                        float* y_out, int64_t batch, int64_t window, int64_t obs, int64_t n_lon, int64_t n_lat,
                        int64_t f_total, int64_t f_used, int flat, void* stream);
/* f4  streaming forecast metrics (scripts/predict.py:53-124, StreamingMetrics.update) for a batch:
 *   y_true, y_pred [B, G, cols]; out float64 [B, cols, 3] = {sum_g err^2, sum_g |err|, spatial anomaly correlation
 *   <yt - mean, yp - mean> / (|yt - mean| |yp - mean| + 1e-8)}; float64 accumulation, fixed summation order. */
size_t gcl_forecast_metrics_workspace_bytes(int64_t batch, int64_t cols);
int gcl_forecast_metrics_f32(const float* y_true, const float* y_pred, double* out, int64_t batch, int64_t n_grid,
                             int64_t cols, void* workspace, size_t workspace_bytes, void* stream);

/* torch.optim.Adam (main.py:212; betas/eps defaults, weight_decay 0) over one flat fp32 buffer.
 * step_count: device int32[1], incremented by the kernel (graph-replay safe). grad_scale multiplies g. */
int gcl_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, float grad_scale,
                 int32_t* step_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCL_B200_H_ */

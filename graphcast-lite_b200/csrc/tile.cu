// Tiled segmented SpMM / single-head GAT aggregation (K2/K3/K4/K6, second generation).
//
// Same arithmetic as spmm.cu / gat.cu (out[b,i,:] = epi(sum_k w[k] x[b,col[k],:]), entries of a row in ascending
// CSR order, no atomics), but a CTA owns a *tile* of rows and the union of their source rows is staged in shared
// memory by asynchronous bulk copies (tile.cuh).  Replaces PyG's propagate (index_select -> mul -> scatter_add_)
// at /root/reference/src/models.py:414 (SimpleConv), :419 (GCNConv) and, with MODE 2, the whole message passing of
// a single-head GATConv (:425): logits, LeakyReLU, segment softmax and aggregation in one kernel.
//
//   gcl_tile_plan_host   host-side greedy tiling of a CSR (once per graph / orientation, cached by the caller)
//   gcl_spmm_tiled_f32   GCN / mean / unit weights (+ bias, PReLU, pre-activation copy): persistent, warp-specialised
//                        CTAs walk the (tile, sample block) items through a 3-stage shared-memory ring
//                        (spmm_ws_kernel: a producer warp issues the copies, 16 consumer warps reduce); heavy
//                        rows (more entries than a tile's union may hold, e.g. the 687-entry polar rows of the
//                        512x256 grid->mesh graph) by a CTA-per-row kernel with a fixed-order two-phase reduction
//   gcl_gat_fwd_tiled_f32  heads == 1 GATConv forward: alpha computed inside the tile CTA
#include <algorithm>
#include <vector>

#include "tile.cuh"
#include "ws.cuh"

namespace gcl {
namespace {

__device__ __forceinline__ float leaky_f(float v, float slope) { return v > 0.f ? v : slope * v; }

// shared-memory carve-up (bytes) for a plan / channel count / SB / number of per-entry weight planes
struct TileSmem {
  size_t xs, ew, re, rid, as, eli, total;
};
inline TileSmem tile_smem(const TileArgs& p, int C, int SB, int wplanes, bool scores) {
  TileSmem s;
  size_t o = 128;                                         // mbarrier + padding
  s.xs = o;  o += (size_t)SB * p.max_union * C * 4;
  s.ew = o;  o += (size_t)wplanes * p.max_entries * 4;
  s.re = o;  o += ((size_t)p.max_rows + 1) * 4;
  s.rid = o; o += (size_t)p.max_rows * 4;
  s.as = o;  o += scores ? (size_t)SB * p.max_union * 4 : 0;
  s.eli = o; o += (size_t)p.max_entries * 2;
  s.total = (o + 15) & ~size_t(15);
  return s;
}

// MODE 0: one weight per CSR entry shared by all samples (w nullable = 1)           -- GCNConv / SimpleConv
// MODE 1: per-sample weights w[b][k] (w_bstride apart) times w_scale                 -- aggregation with given alpha
// MODE 2: weights = segment softmax of LeakyReLU(a_src[col] + a_dst[row]), written to alpha_csr (/ alpha_pyg)
template <int L, int SB, int MODE>
__global__ void __launch_bounds__(kTileThreads)
    spmm_tile_kernel(TileArgs p, TileSmem sm, const float* __restrict__ w, const float* __restrict__ x,
                     float* __restrict__ out, int C, int64_t x_bstride, int64_t out_bstride, int B,
                     const float* __restrict__ bias, const float* __restrict__ prelu_slope, float* __restrict__ z_out,
                     int64_t w_bstride, float w_scale, const float* __restrict__ a_src, const float* __restrict__ a_dst,
                     int64_t a_bstride, float neg_slope, float* __restrict__ alpha_csr, float* __restrict__ alpha_pyg,
                     const int32_t* __restrict__ perm, int64_t nnz) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int WP = MODE == 0 ? 1 : SB;
  float* xs = reinterpret_cast<float*>(smem + sm.xs);
  float* e_w = reinterpret_cast<float*>(smem + sm.ew);
  int32_t* r_e = reinterpret_cast<int32_t*>(smem + sm.re);
  int32_t* r_id = reinterpret_cast<int32_t*>(smem + sm.rid);
  float* as_s = reinterpret_cast<float*>(smem + sm.as);
  uint16_t* e_li = reinterpret_cast<uint16_t*>(smem + sm.eli);
  const uint32_t bar = tile_smem_u32(smem);
  const int tid = threadIdx.x;
  const int b0 = blockIdx.y * SB;
  const int nb = min(SB, B - b0);
  const TileHdr h = tile_header(p, blockIdx.x);

  if (tid == 0) tile_mbar_init(bar, 1);
  __syncthreads();
  tile_issue_rows(p, h, x, x_bstride, C, b0, nb, xs, bar);

  // tile metadata -> shared memory while the rows are in flight
  for (int i = tid; i <= h.nr; i += kTileThreads) {
    r_e[i] = __ldg(p.eptr + h.r0 + i) - h.e0;
    if (i < h.nr) r_id[i] = __ldg(p.rows + h.r0 + i);
  }
  for (int le = tid; le < h.ne; le += kTileThreads) {
    e_li[le] = p.lidx[h.e0 + le];
    if (MODE != 2) {
      const int32_t k = __ldg(p.ek + h.e0 + le);
      if (MODE == 0) {
        e_w[le] = w ? __ldg(w + k) : 1.f;
      } else {
#pragma unroll
        for (int s = 0; s < SB; ++s)
          if (s < nb) e_w[s * p.max_entries + le] = w[(int64_t)(b0 + s) * w_bstride + k] * w_scale;
      }
    }
  }
  if (MODE == 2) {
    for (int idx = tid; idx < h.nu * nb; idx += kTileThreads) {
      const int s = idx / h.nu, u = idx - s * h.nu;
      as_s[s * p.max_union + u] = __ldg(a_src + (int64_t)(b0 + s) * a_bstride + __ldg(p.usrc + h.u0 + u));
    }
  }
  __syncthreads();
  if (MODE == 2) {
    // attention coefficients: one thread per (row, sample); PyG softmax: exp(e - max) / (sum + 1e-16)
    for (int item = tid; item < h.nr * nb; item += kTileThreads) {
      const int s = item / h.nr, i = item - s * h.nr;
      const float adi = __ldg(a_dst + (int64_t)(b0 + s) * a_bstride + r_id[i]);
      const float* asb = as_s + s * p.max_union;
      float* ew = e_w + s * p.max_entries;
      const int le0 = r_e[i], le1 = r_e[i + 1];
      float m = -INFINITY;
      for (int le = le0; le < le1; ++le) {
        const float e = leaky_f(asb[e_li[le]] + adi, neg_slope);
        ew[le] = e;
        m = fmaxf(m, e);
      }
      float l = 0.f;
      for (int le = le0; le < le1; ++le) {
        const float pe = __expf(ew[le] - m);
        ew[le] = pe;
        l += pe;
      }
      const float rl = 1.f / (l + 1e-16f);
      float* ac = alpha_csr + (int64_t)(b0 + s) * nnz;
      float* ap = alpha_pyg ? alpha_pyg + (int64_t)(b0 + s) * nnz : nullptr;
      for (int le = le0; le < le1; ++le) {
        const float al = ew[le] * rl;
        ew[le] = al;
        const int32_t k = __ldg(p.ek + h.e0 + le);
        ac[k] = al;
        if (ap) ap[__ldg(perm + k)] = al;
      }
    }
    __syncthreads();
  }
  tile_mbar_wait(bar, 0);

  // aggregation: a group of L lanes (one 128-bit word each) per row, rows of the tile round-robin over the groups
  constexpr int kGroups = kTileThreads / L;
  const int gl = tid & (L - 1), grp = tid / L;
  const int off = gl * 4;
  const bool live = off < C;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias && live) bv = ldg4(bias + off);
  const float slope = prelu_slope ? __ldg(prelu_slope) : 0.f;
  const int xs_sstride = p.max_union * C;
  for (int i = grp; i < h.nr; i += kGroups) {
    if (!live) continue;
    const int le0 = r_e[i], le1 = r_e[i + 1];
    float4 acc[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int le = le0; le < le1; ++le) {
      const uint32_t li = e_li[le];
      if (li != kMasked) {
        const float* xr = xs + li * C + off;
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          if (s < nb) {
            const float4 v = *reinterpret_cast<const float4*>(xr + s * xs_sstride);
            fma4(acc[s], e_w[(WP == 1 ? 0 : s * p.max_entries) + le], v);
          }
        }
      }
    }
    const int64_t row = r_id[i];
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      if (s >= nb) break;
      float4 a = make_float4(acc[s].x + bv.x, acc[s].y + bv.y, acc[s].z + bv.z, acc[s].w + bv.w);
      const int64_t o = (int64_t)(b0 + s) * out_bstride + row * C + off;
      if (z_out) st4(z_out + o, a);
      if (prelu_slope) a = make_float4(prelu_f(a.x, slope), prelu_f(a.y, slope), prelu_f(a.z, slope), prelu_f(a.w, slope));
      st4(out + o, a);
    }
  }
}

// Rows with more entries than a tile may hold: one CTA of 32 warps per (row, sample).  A group of GS lanes (one
// 128-bit word each, GS = 4..32 >= words per row) takes entries beg + g, beg + g + n_groups, ... (a fixed
// assignment), the partial rows are summed in group order -> deterministic, no atomics.
constexpr int kHeavyThreads = 1024;
template <int GS>
__global__ void __launch_bounds__(kHeavyThreads)
    spmm_heavy_kernel(const int32_t* __restrict__ heavy_rows, const int32_t* __restrict__ rowptr,
                      const int32_t* __restrict__ col, const float* __restrict__ w, const float* __restrict__ x,
                      float* __restrict__ out, int64_t n_in, int C, int64_t x_bstride, int64_t out_bstride,
                      const float* __restrict__ bias, const float* __restrict__ prelu_slope, float* __restrict__ z_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NG = kHeavyThreads / GS;
  float4* part = reinterpret_cast<float4*>(smem);          // [NG][words]
  const int words = C >> 2;
  const int gl = threadIdx.x & (GS - 1), g = threadIdx.x / GS;
  const int64_t row = heavy_rows[blockIdx.x];
  const int b = blockIdx.y;
  const int32_t beg = rowptr[row], end = rowptr[row + 1];
  const float* xb = x + (int64_t)b * x_bstride + gl * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gl < words) {
#pragma unroll 4
    for (int32_t k = beg + g; k < end; k += NG) {
      const int32_t c = __ldg(col + k);
      if (c < n_in) fma4(acc, w ? __ldg(w + k) : 1.f, ldg4(xb + (int64_t)c * C));
    }
    part[g * words + gl] = acc;
  }
  __syncthreads();
  const float slope = prelu_slope ? __ldg(prelu_slope) : 0.f;
  if (threadIdx.x < words) {
    const int wd = threadIdx.x;
    float4 a = part[wd];
    for (int q = 1; q < NG; ++q) {
      const float4 t = part[q * words + wd];
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    if (bias) {
      const float4 bv = ldg4(bias + wd * 4);
      a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
    }
    const int64_t o = (int64_t)b * out_bstride + row * C + wd * 4;
    if (z_out) st4(z_out + o, a);
    if (prelu_slope) a = make_float4(prelu_f(a.x, slope), prelu_f(a.y, slope), prelu_f(a.z, slope), prelu_f(a.w, slope));
    st4(out + o, a);
  }
}

// bf16 rows: a lane owns one 16-byte word = 8 channels (fp32 accumulation, bf16 result)
template <int GS>
__global__ void __launch_bounds__(kHeavyThreads)
    spmm_heavy_bf16_kernel(const int32_t* __restrict__ heavy_rows, const int32_t* __restrict__ rowptr,
                           const int32_t* __restrict__ col, const float* __restrict__ w,
                           const unsigned char* __restrict__ x, unsigned char* __restrict__ out, int64_t n_in, int C,
                           int64_t x_bstride, int64_t out_bstride, const float* __restrict__ bias,
                           const float* __restrict__ prelu_slope, unsigned char* __restrict__ z_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NG = kHeavyThreads / GS;
  float4* part = reinterpret_cast<float4*>(smem);          // [NG][words][2]
  const int words = C >> 3;
  const int gl = threadIdx.x & (GS - 1), g = threadIdx.x / GS;
  const int64_t row = heavy_rows[blockIdx.x];
  const int b = blockIdx.y;
  const int32_t beg = rowptr[row], end = rowptr[row + 1];
  const unsigned char* xb = x + ((int64_t)b * x_bstride) * 2 + gl * 16;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  if (gl < words) {
#pragma unroll 4
    for (int32_t k = beg + g; k < end; k += NG) {
      const int32_t c = __ldg(col + k);
      if (c < n_in) {
        const float wk = w ? __ldg(w + k) : 1.f;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)c * C * 2));
        fma4(a0, wk, bf16x4_lo(u));
        fma4(a1, wk, bf16x4_hi(u));
      }
    }
    part[(g * words + gl) * 2] = a0;
    part[(g * words + gl) * 2 + 1] = a1;
  }
  __syncthreads();
  const float slope = prelu_slope ? __ldg(prelu_slope) : 0.f;
  if (threadIdx.x < words) {
    const int wd = threadIdx.x;
    float4 s0 = part[wd * 2], s1 = part[wd * 2 + 1];
    for (int q = 1; q < NG; ++q) {
      const float4 t0 = part[(q * words + wd) * 2], t1 = part[(q * words + wd) * 2 + 1];
      s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
      s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
    }
    if (bias) {
      const float4 b0 = ldg4(bias + wd * 8), b1 = ldg4(bias + wd * 8 + 4);
      s0.x += b0.x; s0.y += b0.y; s0.z += b0.z; s0.w += b0.w;
      s1.x += b1.x; s1.y += b1.y; s1.z += b1.z; s1.w += b1.w;
    }
    const int64_t o = ((int64_t)b * out_bstride + row * C) * 2 + wd * 16;
    if (z_out) *reinterpret_cast<uint4*>(z_out + o) = pack_bf16x8(s0, s1);
    if (prelu_slope) {
      s0 = ws_prelu4(s0, slope);
      s1 = ws_prelu4(s1, slope);
    }
    *reinterpret_cast<uint4*>(out + o) = pack_bf16x8(s0, s1);
  }
}

// samples per CTA: as many as keep the staged rows within ~64 KB (three CTAs per SM) and <= 4
inline int tile_pick_sb(const TileArgs& p, int64_t C, int64_t B) {
  static const int cap = getenv("GCL_TILE_SB") ? atoi(getenv("GCL_TILE_SB")) : 4;
  static const int64_t budget = getenv("GCL_TILE_SMEM_KB") ? atoll(getenv("GCL_TILE_SMEM_KB")) << 10 : (64 << 10);
  int sb = cap;
  while (sb > 1 && (sb > B || (int64_t)sb * p.max_union * C * 4 > budget)) sb >>= 1;
  return sb;
}

struct TileCall {
  const float *w, *x;
  float* out;
  int C;
  int64_t xbs, obs;
  int B;
  const float *bias, *slope;
  float* z_out;
  int64_t wbs;
  float wsc;
  const float *a_src, *a_dst;
  int64_t abs_;
  float neg;
  float *alpha_csr, *alpha_pyg;
  const int32_t* perm;
  int64_t nnz;
};

template <int L, int SB, int MODE>
int tile_launch(const gcl_tile_plan* plan, const TileCall& c, cudaStream_t s, const char* what) {
  const TileArgs p = tile_args(plan);
  const TileSmem sm = tile_smem(p, c.C, SB, MODE == 0 ? 1 : SB, MODE == 2);
  if (sm.total > 227 * 1024) {
    set_error("%s: tile needs %zu bytes of shared memory", what, sm.total);
    return GCL_ERR_UNSUPPORTED;
  }
  auto kern = spmm_tile_kernel<L, SB, MODE>;
  static bool attr_set = false;               // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, what);
    attr_set = true;
  }
  dim3 grid((unsigned)plan->n_tiles, (unsigned)ceil_div(c.B, SB));
  kern<<<grid, kTileThreads, sm.total, s>>>(p, sm, c.w, c.x, c.out, c.C, c.xbs, c.obs, c.B, c.bias, c.slope, c.z_out,
                                            c.wbs, c.wsc, c.a_src, c.a_dst, c.abs_, c.neg, c.alpha_csr, c.alpha_pyg,
                                            c.perm, c.nnz);
  GCL_CHECK_LAUNCH(what);
  return GCL_OK;
}

template <int L, int MODE>
int tile_dispatch_sb(int sb, const gcl_tile_plan* plan, const TileCall& c, cudaStream_t s, const char* what) {
  if (sb >= 4) return tile_launch<L, 4, MODE>(plan, c, s, what);
  if (sb >= 2) return tile_launch<L, 2, MODE>(plan, c, s, what);
  return tile_launch<L, 1, MODE>(plan, c, s, what);
}

template <int MODE>
int tile_dispatch(const gcl_tile_plan* plan, const TileCall& c, cudaStream_t s, const char* what) {
  const int sb = tile_pick_sb(tile_args(plan), c.C, c.B);
  const int words = c.C / 4;
  if (words <= 4) return tile_dispatch_sb<4, MODE>(sb, plan, c, s, what);
  if (words <= 8) return tile_dispatch_sb<8, MODE>(sb, plan, c, s, what);
  if (words <= 16) return tile_dispatch_sb<16, MODE>(sb, plan, c, s, what);
  return tile_dispatch_sb<32, MODE>(sb, plan, c, s, what);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_plan(const gcl_tile_plan* plan, const char* what) {
  GCL_CHECK_ARG(plan, "%s: null plan", what);
  GCL_CHECK_ARG(plan->n_tiles == 0 || (plan->tile_rowptr && plan->tile_uptr && plan->rows && plan->eptr && plan->ek &&
                                       plan->usrc && plan->lidx),
                "%s: plan with null arrays", what);
  GCL_CHECK_ARG(plan->n_tiles >= 0 && plan->n_heavy >= 0 && plan->max_rows >= 0 && plan->max_union >= 0 &&
                    plan->max_entries >= 0 && plan->max_union < 0xFFFF,
                "%s: bad plan sizes", what);
  return GCL_OK;
}

}  // namespace
}  // namespace gcl

using namespace gcl;

extern "C" int gcl_tile_plan_host(const int32_t* rowptr, const int32_t* col, const int32_t* order, int64_t n_rows,
                                  int64_t n_rows_out, int64_t n_rows_in, int32_t max_rows, int32_t max_union,
                                  int32_t max_entries, int32_t pad_entries, int32_t* tile_rowptr, int32_t* tile_uptr,
                                  int32_t* rows, int32_t* eptr, uint16_t* lidx, int32_t* ek, int32_t* usrc,
                                  int32_t* heavy_rows, int32_t* tile_desc, int64_t* counts) {
  GCL_CHECK_ARG(rowptr && col && tile_rowptr && tile_uptr && rows && eptr && lidx && ek && usrc && heavy_rows &&
                    tile_desc && counts,
                "gcl_tile_plan_host: null pointer argument");
  GCL_CHECK_ARG(n_rows >= 0 && n_rows_out >= 0 && n_rows_out <= n_rows && n_rows_in > 0 && max_rows > 0 &&
                    max_union > 0 && max_union < 0xFFFF && max_entries >= max_union + 3 &&
                    (pad_entries == 1 || pad_entries == 2 || pad_entries == 4),
                "gcl_tile_plan_host: bad sizes");
  const size_t ncol = (size_t)std::max<int64_t>(n_rows_in, 1);
  // stamp[c] == T: column c is in the current tile's union, numbered local[c]; rstamp marks the columns of the row
  // under consideration (distinct count)
  std::vector<int64_t> stamp(ncol, -1), rstamp(ncol, -1);
  std::vector<int32_t> local(ncol, 0);
  int64_t T = 0, nplan = 0, nu_total = 0, ne_total = 0, nheavy = 0;
  int cur_rows = 0, cur_u = 0, cur_e = 0, mr = 0, mu = 0, me = 0;
  tile_rowptr[0] = 0;
  tile_uptr[0] = 0;
  eptr[0] = 0;
  auto close_tile = [&]() {
    if (cur_rows == 0) return;
    mr = std::max(mr, cur_rows); mu = std::max(mu, cur_u); me = std::max(me, cur_e);
    int32_t* d = tile_desc + 8 * T;
    d[0] = tile_rowptr[T]; d[1] = cur_rows; d[2] = tile_uptr[T]; d[3] = cur_u;
    d[4] = eptr[tile_rowptr[T]]; d[5] = cur_e; d[6] = d[7] = 0;
    ++T;                                        // stamps of the closed tile become stale
    tile_rowptr[T] = (int32_t)nplan;
    tile_uptr[T] = (int32_t)nu_total;
    cur_rows = cur_u = cur_e = 0;
  };
  for (int64_t q = 0; q < n_rows; ++q) {
    const int64_t r = order ? order[q] : q;
    GCL_CHECK_ARG(r >= 0 && r < n_rows, "gcl_tile_plan_host: order[%lld] = %lld out of range", (long long)q, (long long)r);
    if (r >= n_rows_out) continue;
    const int32_t beg = rowptr[r], end = rowptr[r + 1];
    const int len = end - beg;
    const int plen = (len + pad_entries - 1) / pad_entries * pad_entries;     // padded entry count of the row
    int distinct = 0, add = 0;                  // distinct sources of the row / those not yet in the tile's union
    for (int32_t k = beg; k < end; ++k) {
      const int32_t c = col[k];
      if (c >= n_rows_in || rstamp[c] == q) continue;
      rstamp[c] = q;
      ++distinct;
      add += stamp[c] != T;
    }
    if (distinct > max_union || plen > max_entries) {
      heavy_rows[nheavy++] = (int32_t)r;
      continue;
    }
    if (cur_rows == max_rows || cur_u + add > max_union || cur_e + plen > max_entries) close_tile();
    rows[nplan] = (int32_t)r;
    for (int32_t k = beg; k < end; ++k) {
      const int32_t c = col[k];
      uint16_t li = kMasked;
      if (c < n_rows_in) {
        if (stamp[c] != T) {
          stamp[c] = T;
          local[c] = cur_u++;
          usrc[nu_total++] = c;
        }
        li = (uint16_t)local[c];
      }
      lidx[ne_total] = li;
      ek[ne_total] = k;
      ++ne_total;
    }
    for (int q2 = len; q2 < plen; ++q2) {       // padding: counts as a zero row with weight 0
      lidx[ne_total] = kMasked;
      ek[ne_total] = -1;
      ++ne_total;
    }
    cur_e += plen;
    ++cur_rows;
    ++nplan;
    eptr[nplan] = (int32_t)ne_total;
  }
  close_tile();
  counts[0] = T; counts[1] = nplan; counts[2] = nu_total; counts[3] = ne_total; counts[4] = nheavy;
  counts[5] = mr; counts[6] = mu; counts[7] = me;
  return GCL_OK;
}

extern "C" int gcl_spmm_tiled_f32(const gcl_tile_plan* plan, const int32_t* ent, const int32_t* rowptr,
                                  const int32_t* col, const float* w, const float* x, float* out, int64_t batch,
                                  int64_t n_rows_in, int64_t channels, int64_t x_bstride, int64_t out_bstride,
                                  const float* bias, const float* prelu_slope, float* z_out, void* stream) {
  if (int rc = check_plan(plan, "gcl_spmm_tiled_f32")) return rc;
  GCL_CHECK_ARG(rowptr && col && x && out && x != out, "gcl_spmm_tiled_f32: null or aliased pointer argument");
  GCL_CHECK_ARG(plan->n_tiles == 0 || (ent && plan->tile_desc && al16(ent) && al16(plan->tile_desc)),
                "gcl_spmm_tiled_f32: the packed entries / tile descriptors are missing or not 16-byte aligned");
  GCL_CHECK_ARG(plan->pad_entries == 2 || plan->pad_entries == 4 || plan->n_tiles == 0,
                "gcl_spmm_tiled_f32: the plan must be built with pad_entries = 2 or 4");
  GCL_CHECK_ARG(channels > 0 && channels % 4 == 0 && channels <= 128 && x_bstride % 4 == 0 && out_bstride % 4 == 0 &&
                    al16(x) && al16(out) && (!bias || al16(bias)) && (!z_out || al16(z_out)),
                "gcl_spmm_tiled_f32: needs 16-byte aligned rows of 4..128 channels (multiple of 4); use gcl_spmm_f32");
  GCL_CHECK_ARG(batch >= 0 && batch <= 65535 && n_rows_in > 0, "gcl_spmm_tiled_f32: bad sizes");
  GCL_CHECK_ARG(n_rows_in * channels < (1ll << 31), "gcl_spmm_tiled_f32: one sample's features must index with 31 bits");
  if (batch == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (plan->n_tiles > 0) {
    WsParams q{};
    q.p = tile_args(plan);
    q.ent = reinterpret_cast<const int2*>(ent);
    q.x = x; q.x_bstride = x_bstride;
    q.out = out; q.out_bstride = out_bstride; q.z_out = z_out; q.bias = bias; q.prelu_slope = prelu_slope;
    q.C = (int)channels; q.B = (int)batch;
    if (int rc = ws_dispatch<0>(q, s, "gcl_spmm_tiled_f32")) return rc;
  }
  if (plan->n_heavy > 0) {
    GCL_CHECK_ARG(plan->heavy_rows, "gcl_spmm_tiled_f32: plan lists heavy rows but has no heavy_rows array");
    dim3 grid((unsigned)plan->n_heavy, (unsigned)batch);
    const int words = (int)channels / 4;
#define GCL_HEAVY(GS)                                                                                            \
  spmm_heavy_kernel<GS><<<grid, kHeavyThreads, (kHeavyThreads / GS) * channels * 4, s>>>(                         \
      plan->heavy_rows, rowptr, col, w, x, out, n_rows_in, (int)channels, x_bstride, out_bstride, bias, prelu_slope, z_out)
    if (words <= 4) GCL_HEAVY(4);
    else if (words <= 8) GCL_HEAVY(8);
    else if (words <= 16) GCL_HEAVY(16);
    else GCL_HEAVY(32);
#undef GCL_HEAVY
    GCL_CHECK_LAUNCH("gcl_spmm_tiled_f32(heavy rows)");
  }
  return GCL_OK;
}

extern "C" int gcl_spmm_tiled_bf16(const gcl_tile_plan* plan, const int32_t* ent, const int32_t* rowptr,
                                  const int32_t* col, const float* w, const void* x, void* out, int64_t batch,
                                  int64_t n_rows_in, int64_t channels, int64_t x_bstride, int64_t out_bstride,
                                  const float* bias, const float* prelu_slope, void* z_out, void* stream) {
  if (int rc = check_plan(plan, "gcl_spmm_tiled_bf16")) return rc;
  GCL_CHECK_ARG(rowptr && col && x && out && x != out, "gcl_spmm_tiled_bf16: null or aliased pointer argument");
  GCL_CHECK_ARG(plan->n_tiles == 0 || (ent && plan->tile_desc && al16(ent) && al16(plan->tile_desc)),
                "gcl_spmm_tiled_bf16: the packed entries / tile descriptors are missing or not 16-byte aligned");
  GCL_CHECK_ARG(plan->pad_entries == 2 || plan->pad_entries == 4 || plan->n_tiles == 0,
                "gcl_spmm_tiled_bf16: the plan must be built with pad_entries = 2 or 4");
  GCL_CHECK_ARG(channels > 0 && channels % 8 == 0 && channels <= 256 && x_bstride % 8 == 0 && out_bstride % 8 == 0 &&
                    al16(x) && al16(out) && (!bias || al16(bias)) && (!z_out || al16(z_out)),
                "gcl_spmm_tiled_bf16: needs 16-byte aligned rows of 8..256 bf16 channels (multiple of 8)");
  GCL_CHECK_ARG(batch >= 0 && batch <= 65535 && n_rows_in > 0 && n_rows_in * channels < (1ll << 31),
                "gcl_spmm_tiled_bf16: bad sizes");
  if (batch == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (plan->n_tiles > 0) {
    WsParams q{};
    q.p = tile_args(plan);
    q.ent = reinterpret_cast<const int2*>(ent);
    q.x = static_cast<const float*>(x); q.x_bstride = x_bstride;          // element strides; the kernel scales by 2 bytes
    q.out = static_cast<float*>(out); q.out_bstride = out_bstride; q.z_out = static_cast<float*>(z_out);
    q.bias = bias; q.prelu_slope = prelu_slope;
    q.C = (int)channels; q.B = (int)batch;
    if (int rc = ws_dispatch_bf16(q, s, "gcl_spmm_tiled_bf16")) return rc;
  }
  if (plan->n_heavy > 0) {
    GCL_CHECK_ARG(plan->heavy_rows, "gcl_spmm_tiled_bf16: plan lists heavy rows but has no heavy_rows array");
    dim3 grid((unsigned)plan->n_heavy, (unsigned)batch);
    const int words = (int)channels / 8;
#define GCL_HEAVY16(GS)                                                                                          \
  spmm_heavy_bf16_kernel<GS><<<grid, kHeavyThreads, (kHeavyThreads / GS) * words * 32, s>>>(                      \
      plan->heavy_rows, rowptr, col, w, static_cast<const unsigned char*>(x), static_cast<unsigned char*>(out),   \
      n_rows_in, (int)channels, x_bstride, out_bstride, bias, prelu_slope, static_cast<unsigned char*>(z_out))
    if (words <= 4) GCL_HEAVY16(4);
    else if (words <= 8) GCL_HEAVY16(8);
    else if (words <= 16) GCL_HEAVY16(16);
    else GCL_HEAVY16(32);
#undef GCL_HEAVY16
    GCL_CHECK_LAUNCH("gcl_spmm_tiled_bf16(heavy rows)");
  }
  return GCL_OK;
}

extern "C" int gcl_gat_fwd_tiled_f32(const gcl_tile_plan* plan, const int32_t* perm, const float* z, const float* a_src,
                                     const float* a_dst, const float* bias, float* out, float* alpha_csr,
                                     float* alpha_pyg, const float* prelu_slope, float* z_out, int64_t batch,
                                     int64_t n_nodes, int64_t nnz, int64_t c, float negative_slope, void* stream) {
  if (int rc = check_plan(plan, "gcl_gat_fwd_tiled_f32")) return rc;
  GCL_CHECK_ARG(z && a_src && a_dst && out && alpha_csr && z != out, "gcl_gat_fwd_tiled_f32: null pointer argument");
  GCL_CHECK_ARG(!alpha_pyg || perm, "gcl_gat_fwd_tiled_f32: alpha_pyg needs perm");
  GCL_CHECK_ARG(plan->n_heavy == 0, "gcl_gat_fwd_tiled_f32: plan has heavy rows; use gcl_gat_fwd_f32");
  GCL_CHECK_ARG(c > 0 && c % 4 == 0 && c <= 128 && al16(z) && al16(out) && (!bias || al16(bias)) &&
                    (!z_out || al16(z_out)),
                "gcl_gat_fwd_tiled_f32: needs 16-byte aligned rows of 4..128 channels (multiple of 4)");
  GCL_CHECK_ARG(batch >= 0 && batch <= 65535 && n_nodes >= 0 && nnz >= 0, "gcl_gat_fwd_tiled_f32: bad sizes");
  if (batch == 0 || n_nodes == 0 || plan->n_tiles == 0) return GCL_OK;
  TileCall cl{nullptr, z, out, (int)c, n_nodes * c, n_nodes * c, (int)batch, bias, prelu_slope, z_out, 0, 1.f,
              a_src, a_dst, n_nodes, negative_slope, alpha_csr, alpha_pyg, perm, nnz};
  return tile_dispatch<2>(plan, cl, static_cast<cudaStream_t>(stream), "gcl_gat_fwd_tiled_f32");
}

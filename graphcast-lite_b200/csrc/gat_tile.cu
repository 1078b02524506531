// K5 on tile plans: backward of the single-head GATConv message passing (models.py:336-357 via PyG autograd),
// fp32, no atomics.  Same two passes and the same per-row summation order as gat.cu, but every feature row a CTA
// gathers is staged in shared memory by asynchronous bulk copies (tile.cuh), so the passes are no longer limited by
// how many gathers a thread can keep in registers:
//   pass 1, receiver-grouped plan: z rows of the tile's union + the tile's dout rows in shared memory;
//           dalpha_k = <dout_i, z_col_k> by a butterfly transpose-reduce, softmax backward, LeakyReLU',
//           g_csr[b,k] and da_dst[b,i] = sum_k g_k
//   pass 2, sender-grouped plan: dout rows of the union in shared memory;
//           da_src[b,j] = sum_k g_k,  dz[b,j,:] = sum_k alpha_k dout[b,i_k,:] + da_src att_src + da_dst att_dst
#include "tile.cuh"

namespace gcl {
namespace {

struct BwdSmem {
  size_t zs, ds, eal, eas, re, rid, eli, total;
};
// pass 1: zs = union z rows, ds = tile dout rows, eal = alpha per (sample, entry), eas = a_src of the union
// pass 2: zs = union dout rows, ds unused, eal = alpha, eas = g per (sample, entry)
inline BwdSmem bwd_smem(const TileArgs& p, int C, int SB, int pass) {
  BwdSmem s;
  size_t o = 128;
  s.zs = o;  o += (size_t)SB * p.max_union * C * 4;
  s.ds = o;  o += pass == 1 ? (size_t)SB * p.max_rows * C * 4 : 0;
  s.eal = o; o += (size_t)SB * p.max_entries * 4;
  s.eas = o; o += pass == 1 ? (size_t)SB * p.max_union * 4 : (size_t)SB * p.max_entries * 4;
  s.re = o;  o += ((size_t)p.max_rows + 1) * 4;
  s.rid = o; o += (size_t)p.max_rows * 4;
  s.eli = o; o += (size_t)p.max_entries * 2;
  s.total = (o + 15) & ~size_t(15);
  return s;
}

template <int L, int SB>
__global__ void __launch_bounds__(kTileThreads)
    gat_bwd_dst_tile_kernel(TileArgs p, BwdSmem sm, const float* __restrict__ z, const float* __restrict__ a_src,
                            const float* __restrict__ a_dst, const float* __restrict__ alpha_csr,
                            const float* __restrict__ dout, float* __restrict__ g_csr, float* __restrict__ da_dst,
                            int64_t N, int64_t nnz, int B, int C, float slope) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NB = L / SB;                 // neighbours per transpose-reduce batch
  float* zs = reinterpret_cast<float*>(smem + sm.zs);
  float* ds = reinterpret_cast<float*>(smem + sm.ds);
  float* e_al = reinterpret_cast<float*>(smem + sm.eal);
  float* as_s = reinterpret_cast<float*>(smem + sm.eas);
  int32_t* r_e = reinterpret_cast<int32_t*>(smem + sm.re);
  int32_t* r_id = reinterpret_cast<int32_t*>(smem + sm.rid);
  uint16_t* e_li = reinterpret_cast<uint16_t*>(smem + sm.eli);
  const uint32_t bar = tile_smem_u32(smem);
  const int tid = threadIdx.x;
  const int b0 = blockIdx.y * SB;
  const int nb = min(SB, B - b0);
  const TileHdr h = tile_header(p, blockIdx.x);
  const uint32_t row_bytes = (uint32_t)C * 4u;

  if (tid == 0) tile_mbar_init(bar, 1);
  __syncthreads();
  if (tid == 0) tile_mbar_expect(bar, (uint32_t)((h.nu + h.nr) * nb) * row_bytes);
  for (int idx = tid; idx < h.nu * nb; idx += kTileThreads) {
    const int s = idx / h.nu, u = idx - s * h.nu;
    const int32_t src = __ldg(p.usrc + h.u0 + u);
    tile_bulk_row(tile_smem_u32(zs + ((size_t)s * p.max_union + u) * C), z + ((int64_t)(b0 + s) * N + src) * C,
                  row_bytes, bar);
    as_s[s * p.max_union + u] = __ldg(a_src + (int64_t)(b0 + s) * N + src);
  }
  for (int idx = tid; idx < h.nr * nb; idx += kTileThreads) {
    const int s = idx / h.nr, i = idx - s * h.nr;
    const int32_t row = __ldg(p.rows + h.r0 + i);
    tile_bulk_row(tile_smem_u32(ds + ((size_t)s * p.max_rows + i) * C), dout + ((int64_t)(b0 + s) * N + row) * C,
                  row_bytes, bar);
  }
  for (int i = tid; i <= h.nr; i += kTileThreads) {
    r_e[i] = __ldg(p.eptr + h.r0 + i) - h.e0;
    if (i < h.nr) r_id[i] = __ldg(p.rows + h.r0 + i);
  }
  for (int le = tid; le < h.ne; le += kTileThreads) {
    e_li[le] = p.lidx[h.e0 + le];
    const int32_t k = __ldg(p.ek + h.e0 + le);
#pragma unroll
    for (int s = 0; s < SB; ++s)
      if (s < nb) e_al[s * p.max_entries + le] = __ldg(alpha_csr + (int64_t)(b0 + s) * nnz + k);
  }
  __syncthreads();
  tile_mbar_wait(bar, 0);

  constexpr int kGroups = kTileThreads / L;
  const int lane = tid & 31;
  const int gl = tid & (L - 1), grp = tid / L;
  const unsigned mask = tile_group_mask<L>(lane);
  const int off = gl * 4 < C ? gl * 4 : 0;          // idle lanes re-read word 0 (their products are dropped)
  const bool glive = gl * 4 < C;
  const int jj = gl / SB, s_me = gl % SB;           // after the transpose-reduce this lane owns (neighbour jj, sample s_me)
  const bool s_ok = s_me < nb;
  const int64_t bs = b0 + s_me;
  for (int i = grp; i < h.nr; i += kGroups) {
    const int le0 = r_e[i], le1 = r_e[i + 1];
    const int64_t row = r_id[i];
    float4 dv[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      const int sc = min(s, nb - 1);
      dv[s] = glive ? *reinterpret_cast<const float4*>(ds + ((size_t)sc * p.max_rows + i) * C + off)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // finished dalpha of this lane's (edge, sample) in the batch starting at local entry `base`
    auto batch = [&](int base) -> float {
      float pr[L];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int le = min(base + j, le1 - 1);      // past the row end: re-read the last edge, dropped below
        const float* zr = zs + (size_t)e_li[le] * C + off;
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          const int sc = min(s, nb - 1);
          pr[j * SB + s] = dot4(dv[s], *reinterpret_cast<const float4*>(zr + (size_t)sc * p.max_union * C));
        }
      }
      return tile_xreduce<L>(pr, gl, mask);
    };
    const float adst = s_ok ? __ldg(a_dst + bs * N + row) : 0.f;
    const float* al_s = e_al + s_me * p.max_entries;
    const float* as_me = as_s + s_me * p.max_union;
    float tp = 0.f, gs = 0.f;
    if (le1 - le0 <= L) {                           // at most SB batches: dalpha stays in registers
      float dal[SB], al[SB];
#pragma unroll
      for (int q = 0; q < SB; ++q) {
        dal[q] = al[q] = 0.f;
        const int base = le0 + q * NB;
        if (base < le1) {                           // group-uniform
          const float r = batch(base);
          const int le = base + jj;
          if (le < le1 && s_ok) {
            dal[q] = r;
            al[q] = al_s[le];
            tp += al[q] * r;
          }
        }
      }
      const float t = tile_gsum_strided<L, SB>(tp, mask);
#pragma unroll
      for (int q = 0; q < SB; ++q) {
        const int le = le0 + q * NB + jj;
        if (le < le1 && s_ok) {
          const float pre = as_me[e_li[le]] + adst;
          const float g = al[q] * (dal[q] - t) * (pre > 0.f ? 1.f : slope);
          g_csr[bs * nnz + __ldg(p.ek + h.e0 + le)] = g;
          gs += g;
        }
      }
    } else {
      for (int base = le0; base < le1; base += NB) {
        const float r = batch(base);
        const int le = base + jj;
        if (le < le1 && s_ok) {
          tp += al_s[le] * r;
          g_csr[bs * nnz + __ldg(p.ek + h.e0 + le)] = r;     // staged; re-read below by the same lane
        }
      }
      const float t = tile_gsum_strided<L, SB>(tp, mask);
      for (int base = le0; base < le1; base += NB) {
        const int le = base + jj;
        if (le < le1 && s_ok) {
          const int64_t idx = bs * nnz + __ldg(p.ek + h.e0 + le);
          const float pre = as_me[e_li[le]] + adst;
          const float g = al_s[le] * (g_csr[idx] - t) * (pre > 0.f ? 1.f : slope);
          g_csr[idx] = g;
          gs += g;
        }
      }
    }
    gs = tile_gsum_strided<L, SB>(gs, mask);
    if (jj == 0 && s_ok) da_dst[bs * N + row] = gs;
  }
}

template <int L, int SB>
__global__ void __launch_bounds__(kTileThreads)
    gat_bwd_src_tile_kernel(TileArgs p, BwdSmem sm, const int32_t* __restrict__ t2r, const float* __restrict__ alpha_csr,
                            const float* __restrict__ g_csr, const float* __restrict__ att_src,
                            const float* __restrict__ att_dst, const float* __restrict__ dout,
                            const float* __restrict__ da_dst, float* __restrict__ da_src, float* __restrict__ dz,
                            int64_t N, int64_t nnz, int B, int C) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* xs = reinterpret_cast<float*>(smem + sm.zs);
  float* e_al = reinterpret_cast<float*>(smem + sm.eal);
  float* e_g = reinterpret_cast<float*>(smem + sm.eas);
  int32_t* r_e = reinterpret_cast<int32_t*>(smem + sm.re);
  int32_t* r_id = reinterpret_cast<int32_t*>(smem + sm.rid);
  uint16_t* e_li = reinterpret_cast<uint16_t*>(smem + sm.eli);
  const uint32_t bar = tile_smem_u32(smem);
  const int tid = threadIdx.x;
  const int b0 = blockIdx.y * SB;
  const int nb = min(SB, B - b0);
  const TileHdr h = tile_header(p, blockIdx.x);

  if (tid == 0) tile_mbar_init(bar, 1);
  __syncthreads();
  tile_issue_rows(p, h, dout, N * C, C, b0, nb, xs, bar);
  for (int i = tid; i <= h.nr; i += kTileThreads) {
    r_e[i] = __ldg(p.eptr + h.r0 + i) - h.e0;
    if (i < h.nr) r_id[i] = __ldg(p.rows + h.r0 + i);
  }
  for (int le = tid; le < h.ne; le += kTileThreads) {
    e_li[le] = p.lidx[h.e0 + le];
    const int32_t kr = __ldg(t2r + __ldg(p.ek + h.e0 + le));
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      if (s < nb) {
        e_al[s * p.max_entries + le] = __ldg(alpha_csr + (int64_t)(b0 + s) * nnz + kr);
        e_g[s * p.max_entries + le] = g_csr[(int64_t)(b0 + s) * nnz + kr];
      }
    }
  }
  __syncthreads();
  tile_mbar_wait(bar, 0);

  constexpr int kGroups = kTileThreads / L;
  const int gl = tid & (L - 1), grp = tid / L;
  const int off = gl * 4;
  const bool live = off < C;
  float4 as4 = make_float4(0.f, 0.f, 0.f, 0.f), ad4 = as4;
  if (live) {
    as4 = ldg4(att_src + off);
    ad4 = ldg4(att_dst + off);
  }
  const int xs_sstride = p.max_union * C;
  for (int i = grp; i < h.nr; i += kGroups) {
    const int le0 = r_e[i], le1 = r_e[i + 1];
    const int64_t row = r_id[i];
    float4 acc[SB];
    float gsv[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
      gsv[s] = 0.f;
    }
#pragma unroll 4
    for (int le = le0; le < le1; ++le) {
      const float* xr = xs + (size_t)e_li[le] * C + (live ? off : 0);
#pragma unroll
      for (int s = 0; s < SB; ++s) {
        if (s < nb) {
          gsv[s] += e_g[s * p.max_entries + le];            // ascending entry order, every lane the same sum
          fma4(acc[s], e_al[s * p.max_entries + le], *reinterpret_cast<const float4*>(xr + s * xs_sstride));
        }
      }
    }
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      if (s >= nb) break;
      const int64_t node = (int64_t)(b0 + s) * N + row;
      if (gl == 0) da_src[node] = gsv[s];
      if (live) {
        const float dad = __ldg(da_dst + node);
        fma4(acc[s], gsv[s], as4);
        fma4(acc[s], dad, ad4);
        st4(dz + node * C + off, acc[s]);
      }
    }
  }
}

inline int bwd_pick_sb(const TileArgs& p, int64_t C, int64_t B, int rows_too) {
  static const int cap = getenv("GCL_TILE_BWD_SB") ? atoi(getenv("GCL_TILE_BWD_SB")) : 2;
  static const int64_t budget = getenv("GCL_TILE_BWD_SMEM_KB") ? atoll(getenv("GCL_TILE_BWD_SMEM_KB")) << 10 : (100 << 10);
  int sb = cap;
  while (sb > 1 && (sb > B || (int64_t)sb * (p.max_union + (rows_too ? p.max_rows : 0)) * C * 4 > budget)) sb >>= 1;
  return sb;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename K>
int set_smem_attr(K kern, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return fail_cuda(e, what);
  return GCL_OK;
}

}  // namespace
}  // namespace gcl

using namespace gcl;

#define BWD_LAUNCH(KERNEL, LC, SBC, SM, GRID, ...)                                     \
  do {                                                                                 \
    auto kern = KERNEL<LC, SBC>;                                                       \
    static bool attr = false;                                                          \
    if (!attr) {                                                                       \
      if (int rc = set_smem_attr(kern, "gcl_gat_bwd_tiled_f32")) return rc;            \
      attr = true;                                                                     \
    }                                                                                  \
    kern<<<GRID, kTileThreads, (SM).total, s>>>(__VA_ARGS__);                          \
  } while (0)
#define BWD_DISPATCH(KERNEL, LV, SBV, SM, GRID, ...)                                   \
  do {                                                                                 \
    if (SBV == 2) {                                                                    \
      if (LV == 4) BWD_LAUNCH(KERNEL, 4, 2, SM, GRID, __VA_ARGS__);                    \
      else if (LV == 8) BWD_LAUNCH(KERNEL, 8, 2, SM, GRID, __VA_ARGS__);               \
      else if (LV == 16) BWD_LAUNCH(KERNEL, 16, 2, SM, GRID, __VA_ARGS__);             \
      else BWD_LAUNCH(KERNEL, 32, 2, SM, GRID, __VA_ARGS__);                           \
    } else {                                                                           \
      if (LV == 4) BWD_LAUNCH(KERNEL, 4, 1, SM, GRID, __VA_ARGS__);                    \
      else if (LV == 8) BWD_LAUNCH(KERNEL, 8, 1, SM, GRID, __VA_ARGS__);               \
      else if (LV == 16) BWD_LAUNCH(KERNEL, 16, 1, SM, GRID, __VA_ARGS__);             \
      else BWD_LAUNCH(KERNEL, 32, 1, SM, GRID, __VA_ARGS__);                           \
    }                                                                                  \
  } while (0)

extern "C" int gcl_gat_bwd_tiled_f32(const gcl_tile_plan* plan, const gcl_tile_plan* plan_t, const int32_t* t2r,
                                     const float* z, const float* a_src, const float* a_dst, const float* alpha_csr,
                                     const float* att_src, const float* att_dst, const float* dout, float* g_csr,
                                     float* da_src, float* da_dst, float* dz, int64_t batch, int64_t n_nodes,
                                     int64_t nnz, int64_t c, float negative_slope, void* stream) {
  GCL_CHECK_ARG(plan && plan_t && t2r && z && a_src && a_dst && alpha_csr && att_src && att_dst && dout && g_csr &&
                    da_src && da_dst && dz,
                "gcl_gat_bwd_tiled_f32: null pointer argument");
  GCL_CHECK_ARG(plan->n_heavy == 0 && plan_t->n_heavy == 0, "gcl_gat_bwd_tiled_f32: plans with heavy rows; use gcl_gat_bwd_f32");
  GCL_CHECK_ARG(plan->max_union < 0xFFFF && plan_t->max_union < 0xFFFF, "gcl_gat_bwd_tiled_f32: bad plan");
  GCL_CHECK_ARG(c > 0 && c % 4 == 0 && c <= 128 && al16(z) && al16(dout) && al16(dz) && al16(att_src) && al16(att_dst),
                "gcl_gat_bwd_tiled_f32: needs 16-byte aligned rows of 4..128 channels (multiple of 4)");
  GCL_CHECK_ARG(batch >= 0 && batch <= 65535 && n_nodes >= 0 && nnz >= 0, "gcl_gat_bwd_tiled_f32: bad sizes");
  if (batch == 0 || n_nodes == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int words = (int)c / 4;
  const int l = words <= 4 ? 4 : words <= 8 ? 8 : words <= 16 ? 16 : 32;
  {
    const TileArgs p = tile_args(plan);
    const int sb = bwd_pick_sb(p, c, batch, 1);
    const BwdSmem sm = bwd_smem(p, (int)c, sb, 1);
    if (sm.total > 227 * 1024) {
      set_error("gcl_gat_bwd_tiled_f32: pass 1 needs %zu bytes of shared memory", sm.total);
      return GCL_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)plan->n_tiles, (unsigned)ceil_div(batch, sb));
    if (plan->n_tiles > 0)
      BWD_DISPATCH(gat_bwd_dst_tile_kernel, l, sb, sm, grid, p, sm, z, a_src, a_dst, alpha_csr, dout, g_csr, da_dst,
                   n_nodes, nnz, (int)batch, (int)c, negative_slope);
    GCL_CHECK_LAUNCH("gcl_gat_bwd_tiled_f32(dst pass)");
  }
  {
    const TileArgs p = tile_args(plan_t);
    const int sb = bwd_pick_sb(p, c, batch, 0);
    const BwdSmem sm = bwd_smem(p, (int)c, sb, 2);
    if (sm.total > 227 * 1024) {
      set_error("gcl_gat_bwd_tiled_f32: pass 2 needs %zu bytes of shared memory", sm.total);
      return GCL_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)plan_t->n_tiles, (unsigned)ceil_div(batch, sb));
    if (plan_t->n_tiles > 0)
      BWD_DISPATCH(gat_bwd_src_tile_kernel, l, sb, sm, grid, p, sm, t2r, alpha_csr, g_csr, att_src, att_dst, dout,
                   da_dst, da_src, dz, n_nodes, nnz, (int)batch, (int)c);
    GCL_CHECK_LAUNCH("gcl_gat_bwd_tiled_f32(src pass)");
  }
  return GCL_OK;
}

// ================================================================================================================
// Single-head GATConv on the persistent warp-specialised engine (ws.cuh).  The attention coefficients live in PLAN
// order so that a tile's coefficients are one contiguous run the producers can copy:
//   alpha_f [B, Ef]  receiver-grouped plan order      (forward aggregation, backward pass 1)
//   alr_f   [B, Ef]  alpha * LeakyReLU'(logit)         (backward pass 1: g = alr (dalpha - t))
//   alpha_t [B, Et]  sender-grouped plan order         (backward pass 2)
//   g_t     [B, Et]  d(logit), written by pass 1 in sender-grouped plan order through f2t (plan entry -> plan entry)
// Pads of the plans (rows are padded to entry pairs) hold 0 / must be zero-initialised by the caller (alpha_t, g_t).
#include "ws.cuh"

namespace gcl {
namespace {

// Attention coefficients in all three layouts.  8 lanes per (sample, plan row), lane j owns entry j of a chunk of 8
// (every mesh row is one chunk): logits, segment softmax by group shuffles (PyG: exp(e - max) / (sum + 1e-16)),
// coalesced plan-order stores.  pcol = plan-order sender ids (-1 for pads).
__global__ void __launch_bounds__(256)
    gat_alpha_plan_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ eptr,
                          const int32_t* __restrict__ ek, const int32_t* __restrict__ pcol,
                          const int32_t* __restrict__ perm, const int32_t* __restrict__ f2t,
                          const float* __restrict__ a_src, const float* __restrict__ a_dst,
                          float* __restrict__ alpha_f, float* __restrict__ alr_f, float* __restrict__ alpha_t,
                          float* __restrict__ alpha_pyg, int n_plan_rows, int64_t N, int B, int64_t Ef, int64_t Et,
                          int64_t nnz, float slope) {
  const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
  if (t >= (int64_t)B * n_plan_rows) return;                   // whole 8-lane groups leave together
  const int lane = threadIdx.x & 31, gl = lane & 7;
  const unsigned mask = 0xffu << (lane & ~7);
  const int i = (int)(t % n_plan_rows);
  const int64_t b = t / n_plan_rows;
  const int32_t row = __ldg(rows + i);
  const int32_t e0 = __ldg(eptr + i), e1 = __ldg(eptr + i + 1);
  const float adi = a_dst[b * N + row];
  const float* as = a_src + b * N;
  float* af = alpha_f + b * Ef;
  float* ar = alr_f + b * Ef;
  float* at = alpha_t + b * Et;
  float* ap = alpha_pyg ? alpha_pyg + b * nnz : nullptr;
  if (e1 - e0 <= 8) {                                          // one chunk: the lane keeps its entry in registers
    const int32_t e = e0 + gl;
    const int32_t c = e < e1 ? __ldg(pcol + e) : -1;
    const float pre = c >= 0 ? as[c] + adi : 0.f;
    const float lg = c >= 0 ? (pre > 0.f ? pre : slope * pre) : -INFINITY;
    float m = lg;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(mask, m, o, 8));
    const float ex = c >= 0 ? __expf(lg - m) : 0.f;
    float l = ex;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) l += __shfl_xor_sync(mask, l, o, 8);
    const float al = ex * (1.f / (l + 1e-16f));
    if (e < e1) {
      af[e] = al;
      ar[e] = pre > 0.f ? al : al * slope;
      if (c >= 0) {
        at[__ldg(f2t + e)] = al;
        if (ap) ap[__ldg(perm + __ldg(ek + e))] = al;
      }
    }
    return;
  }
  float m = -INFINITY;
  for (int32_t e = e0 + gl; e < e1; e += 8) {
    const int32_t c = __ldg(pcol + e);
    if (c >= 0) {
      const float pre = as[c] + adi;
      m = fmaxf(m, pre > 0.f ? pre : slope * pre);
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(mask, m, o, 8));
  float l = 0.f;
  for (int32_t e = e0 + gl; e < e1; e += 8) {
    const int32_t c = __ldg(pcol + e);
    if (c >= 0) {
      const float pre = as[c] + adi;
      l += __expf((pre > 0.f ? pre : slope * pre) - m);
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) l += __shfl_xor_sync(mask, l, o, 8);
  const float rl = 1.f / (l + 1e-16f);
  for (int32_t e = e0 + gl; e < e1; e += 8) {
    const int32_t c = __ldg(pcol + e);
    float al = 0.f, alr = 0.f;
    if (c >= 0) {
      const float pre = as[c] + adi;
      al = __expf((pre > 0.f ? pre : slope * pre) - m) * rl;
      alr = pre > 0.f ? al : al * slope;
      at[__ldg(f2t + e)] = al;
      if (ap) ap[__ldg(perm + __ldg(ek + e))] = al;
    }
    af[e] = al;
    ar[e] = alr;
  }
}

int check_ws_plan(const gcl_tile_plan* plan, const char* what) {
  GCL_CHECK_ARG(plan && plan->n_tiles > 0 && plan->tile_desc && plan->rows && plan->eptr && plan->ek && plan->usrc,
                "%s: missing plan arrays", what);
  GCL_CHECK_ARG(plan->pad_entries == 2 || plan->pad_entries == 4, "%s: plans must be built with pad_entries = 2", what);
  GCL_CHECK_ARG(plan->n_heavy == 0, "%s: plans with heavy rows are not supported; use gcl_gat_fwd_f32 / gcl_gat_bwd_f32", what);
  return GCL_OK;
}

}  // namespace
}  // namespace gcl

extern "C" int gcl_gat_ws_supported(const gcl_tile_plan* plan, const gcl_tile_plan* plan_t, int64_t c, int64_t batch) {
  if (!plan || !plan_t || plan->n_heavy || plan_t->n_heavy || plan->n_tiles <= 0 || plan_t->n_tiles <= 0) return 0;
  if (plan->pad_entries < 2 || plan_t->pad_entries < 2 || c <= 0 || c % 4 || c > 128 || batch <= 0) return 0;
  return ws_pick_sb(tile_args(plan), (int)c, (int)batch, 1) > 0 && ws_pick_sb(tile_args(plan), (int)c, (int)batch, 3) > 0 &&
         ws_pick_sb(tile_args(plan_t), (int)c, (int)batch, 2) > 0;
}

extern "C" int gcl_gat_fwd_ws_f32(const gcl_tile_plan* plan, const int32_t* ent, const int32_t* pcol, const int32_t* perm,
                                  const int32_t* f2t, const float* z, const float* a_src, const float* a_dst,
                                  const float* bias, float* out, float* alpha_f, float* alr_f, float* alpha_t,
                                  float* alpha_pyg, const float* prelu_slope, float* z_out, int64_t batch,
                                  int64_t n_nodes, int64_t nnz, int64_t c, int64_t ef, int64_t et,
                                  float negative_slope, void* stream) {
  if (int rc = check_ws_plan(plan, "gcl_gat_fwd_ws_f32")) return rc;
  GCL_CHECK_ARG(ent && pcol && f2t && z && a_src && a_dst && out && alpha_f && alr_f && alpha_t && z != out,
                "gcl_gat_fwd_ws_f32: null pointer argument");
  GCL_CHECK_ARG(!alpha_pyg || perm, "gcl_gat_fwd_ws_f32: alpha_pyg needs perm");
  GCL_CHECK_ARG(c > 0 && c % 4 == 0 && c <= 128 && al16(z) && al16(out) && al16(ent) && (!bias || al16(bias)) &&
                    (!z_out || al16(z_out)) && ef % 2 == 0 && et % 2 == 0 && al16(alpha_f),
                "gcl_gat_fwd_ws_f32: needs 16-byte aligned rows of 4..128 channels and even plan entry counts");
  GCL_CHECK_ARG(batch >= 0 && batch <= 65535 && n_nodes > 0 && n_nodes * c < (1ll << 31), "gcl_gat_fwd_ws_f32: bad sizes");
  if (batch == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the plan lists every row exactly once (no heavy rows): n_plan_rows == n_nodes
  gat_alpha_plan_kernel<<<(unsigned)ceil_div(batch * n_nodes * 8, 256), 256, 0, s>>>(
      plan->rows, plan->eptr, plan->ek, pcol, perm, f2t, a_src, a_dst, alpha_f, alr_f, alpha_t, alpha_pyg, (int)n_nodes,
      n_nodes, (int)batch, ef, et, nnz, negative_slope);
  GCL_CHECK_LAUNCH("gcl_gat_fwd_ws_f32(alpha)");
  WsParams q{};
  q.p = tile_args(plan);
  q.ent = reinterpret_cast<const int2*>(ent);
  q.x = z; q.x_bstride = n_nodes * c;
  q.wa = alpha_f; q.w_bstride = ef;
  q.out = out; q.out_bstride = n_nodes * c; q.z_out = z_out; q.bias = bias; q.prelu_slope = prelu_slope;
  q.n_nodes = n_nodes; q.C = (int)c; q.B = (int)batch;
  return ws_dispatch<1>(q, s, "gcl_gat_fwd_ws_f32");
}

extern "C" int gcl_gat_bwd_ws_f32(const gcl_tile_plan* plan, const int32_t* ent, const gcl_tile_plan* plan_t,
                                  const int32_t* ent_t, const int32_t* f2t, const float* z, const float* alpha_f,
                                  const float* alr_f, const float* alpha_t, const float* att_src, const float* att_dst,
                                  const float* dout, float* g_t, float* da_src, float* da_dst, float* dz, int64_t batch,
                                  int64_t n_nodes, int64_t c, int64_t ef, int64_t et, void* stream) {
  if (int rc = check_ws_plan(plan, "gcl_gat_bwd_ws_f32")) return rc;
  if (int rc = check_ws_plan(plan_t, "gcl_gat_bwd_ws_f32")) return rc;
  GCL_CHECK_ARG(ent && ent_t && f2t && z && alpha_f && alr_f && alpha_t && att_src && att_dst && dout && g_t && da_src &&
                    da_dst && dz,
                "gcl_gat_bwd_ws_f32: null pointer argument");
  GCL_CHECK_ARG(c > 0 && c % 4 == 0 && c <= 128 && al16(z) && al16(dout) && al16(dz) && al16(att_src) && al16(att_dst) &&
                    al16(ent) && al16(ent_t) && ef % 2 == 0 && et % 2 == 0,
                "gcl_gat_bwd_ws_f32: needs 16-byte aligned rows of 4..128 channels and even plan entry counts");
  GCL_CHECK_ARG(batch >= 0 && batch <= 65535 && n_nodes > 0 && n_nodes * c < (1ll << 31), "gcl_gat_bwd_ws_f32: bad sizes");
  if (batch == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    WsParams q{};
    q.p = tile_args(plan);
    q.ent = reinterpret_cast<const int2*>(ent);
    q.x = z; q.x2 = dout; q.x_bstride = n_nodes * c;
    q.wa = alpha_f; q.wb = alr_f; q.w_bstride = ef;
    q.g_t = g_t; q.g_bstride = et; q.f2t = f2t; q.da_dst_out = da_dst;
    q.n_nodes = n_nodes; q.C = (int)c; q.B = (int)batch;
    if (int rc = ws_dispatch<3>(q, s, "gcl_gat_bwd_ws_f32(dst pass)")) return rc;
  }
  WsParams q{};
  q.p = tile_args(plan_t);
  q.ent = reinterpret_cast<const int2*>(ent_t);
  q.x = dout; q.x_bstride = n_nodes * c;
  q.wa = alpha_t; q.wb = g_t; q.w_bstride = et;
  q.out = dz; q.out_bstride = n_nodes * c;
  q.att_src = att_src; q.att_dst = att_dst; q.da_dst_in = da_dst; q.da_src_out = da_src;
  q.n_nodes = n_nodes; q.C = (int)c; q.B = (int)batch;
  return ws_dispatch<2>(q, s, "gcl_gat_bwd_ws_f32(src pass)");
}

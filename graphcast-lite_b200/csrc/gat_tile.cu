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

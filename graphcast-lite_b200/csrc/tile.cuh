// Tile plans and the shared-memory staging used by the tiled aggregation kernels (tile.cu, gat_tile.cu).
//
// A *tile* is a set of up to max_rows CSR rows (receivers in the forward plan, senders in the transposed one)
// together with the union of the rows they gather from (<= max_union distinct rows).  One CTA stages the union's
// feature rows of SB samples in shared memory with one asynchronous bulk copy per row (cp.async.bulk + mbarrier
// complete_tx: bytes in flight do not cost registers) and then reduces every row of the tile out of shared
// memory.  A feature row therefore crosses L2 -> SM |union| / |rows| times (1.1 - 2x on the model's graphs)
// instead of once per incident edge (7.4x on the multi-mesh), and never through a register-limited gather.
#pragma once
#include "common.cuh"

namespace gcl {

// device view of a gcl_tile_plan (kernel parameter)
struct TileArgs {
  const int32_t* tile_rowptr;
  const int32_t* tile_uptr;
  const int32_t* rows;
  const int32_t* eptr;
  const int32_t* ek;
  const int32_t* usrc;
  const uint16_t* lidx;
  const int32_t* tile_desc;
  int max_rows, max_union, max_entries, n_tiles;
};

inline TileArgs tile_args(const gcl_tile_plan* p) {
  return TileArgs{p->tile_rowptr, p->tile_uptr, p->rows, p->eptr, p->ek, p->usrc, p->lidx, p->tile_desc,
                  p->max_rows, p->max_union, p->max_entries, p->n_tiles};
}

constexpr int kTileThreads = 256;
constexpr uint16_t kMasked = 0xFFFFu;   // lidx of an entry whose column lies past n_rows_in (counts as a zero row)

__device__ __forceinline__ uint32_t tile_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tile_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tile_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tile_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "TILE_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni TILE_WAIT_DONE;\n\t"
      "bra.uni TILE_WAIT_LOOP;\n\t"
      "TILE_WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// one row: global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tile_bulk_row(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Position of this CTA's tile in the plan
struct TileHdr {
  int r0, nr, u0, nu, e0, ne;
};
__device__ __forceinline__ TileHdr tile_header(const TileArgs& p, int t) {
  TileHdr h;
  h.r0 = __ldg(p.tile_rowptr + t);
  h.nr = __ldg(p.tile_rowptr + t + 1) - h.r0;
  h.u0 = __ldg(p.tile_uptr + t);
  h.nu = __ldg(p.tile_uptr + t + 1) - h.u0;
  h.e0 = __ldg(p.eptr + h.r0);
  h.ne = __ldg(p.eptr + h.r0 + h.nr) - h.e0;
  return h;
}

// Issue the bulk copies of the tile's union rows for samples b0 .. b0+nb-1 of `x` ([B][rows][C], sample stride
// x_bstride) into xs[s][u][C] (u-stride C, sample stride max_union * C).  Called by all threads after the mbarrier
// has been initialised and made visible (__syncthreads).  Thread 0 posts the expected byte count.
__device__ __forceinline__ void tile_issue_rows(const TileArgs& p, const TileHdr& h, const float* __restrict__ x,
                                                int64_t x_bstride, int C, int b0, int nb, float* xs, uint32_t bar) {
  const uint32_t row_bytes = (uint32_t)C * 4u;
  const int total = h.nu * nb;
  if (threadIdx.x == 0) tile_mbar_expect(bar, (uint32_t)total * row_bytes);
  for (int idx = threadIdx.x; idx < total; idx += kTileThreads) {
    const int s = idx / h.nu, u = idx - s * h.nu;
    const int32_t src = __ldg(p.usrc + h.u0 + u);
    tile_bulk_row(tile_smem_u32(xs + ((size_t)s * p.max_union + u) * C),
                  x + (int64_t)(b0 + s) * x_bstride + (int64_t)src * C, row_bytes, bar);
  }
}

}  // namespace gcl

namespace gcl {
// ---- lane-group helpers shared by the tiled attention kernels ------------------------------------------------
template <int L>
__device__ __forceinline__ unsigned tile_group_mask(int lane) {
  return (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << ((lane / L) * L));
}
// Butterfly transpose-reduce: every lane of an L-lane group holds L partial values p[0..L); afterwards lane g holds
// the sum over the group's lanes of p[g].  L - 1 shuffles for L reductions.
template <int L>
__device__ __forceinline__ float tile_xreduce(float (&p)[L], int gl, unsigned mask) {
#pragma unroll
  for (int o = L / 2; o >= 1; o >>= 1) {
    const bool up = (gl & o) != 0;
#pragma unroll
    for (int m = 0; m < o; ++m) {
      const float send = up ? p[m] : p[m + o];
      const float keep = up ? p[m + o] : p[m];
      p[m] = keep + __shfl_xor_sync(mask, send, o, L);
    }
  }
  return p[0];
}
// sum over the lanes of a group that share gl % SB (strides SB, 2 SB, ..., L / 2)
template <int L, int SB>
__device__ __forceinline__ float tile_gsum_strided(float v, unsigned mask) {
#pragma unroll
  for (int o = SB; o < L; o <<= 1) v += __shfl_xor_sync(mask, v, o, L);
  return v;
}
__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {
  a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
}  // namespace gcl

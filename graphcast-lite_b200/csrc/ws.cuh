// Persistent, warp-specialised aggregation over a tile plan (pad_entries = 2): the engine behind gcl_spmm_tiled_f32
// and the single-head GATConv forward / backward (gcl_gat_*_ws_f32).
//
// One CTA per SM walks items (sample block, tile) = blockIdx.x, + gridDim.x, ... through a ring of kWsStages
// shared-memory stages.  A stage holds what one item needs: the union's feature rows of SB samples (+ one all-zero
// row that pad / masked entries point to), the tile's packed entries, row ids, entry offsets, descriptor and --
// depending on MODE -- per-sample per-entry scalars (attention) and the tile's own rows of a second tensor.
//   warps 0..kWsProducerWarps-1 (producers): wait until the consumers have released the stage (mbarrier `empty`),
//     issue the item's copies -- 16-byte cp.async chunks for the rows (source-row ids come from a small shared
//     buffer filled one item ahead, tile descriptors from registers loaded two items ahead), 4/8/16-byte cp.async
//     for the index / scalar data -- and let the copies themselves signal `full`
//     (cp.async.mbarrier.arrive.noinc): producers never wait for data and run up to kWsStages items ahead.
//   the other warps (consumers): wait for `full`, reduce the tile's rows out of shared memory, release the stage.
// History (profiles/README.md): load -> sync -> compute inside one short-lived CTA, and a version in which every
// warp both copied and computed, were bound by the per-item latency chain (descriptor -> source ids -> rows ->
// barrier), not by HBM or the gathers; with one producer warp the consumers starved (57% of samples in the `full`
// wait); 8-12 producer warps feed them.
// The reductions are written for instruction count (ncu: 67 M warp instructions for 10 M FFMAs before): a lane owns
// TWO 128-bit words of a row (a group of L = words/2 lanes per row), entries come as pairs from one LDS.128, pad
// entries point at the zero row (no bounds test in the loop), products use the packed FFMA2 (fma.rn.f32x2,
// bit-identical to two fmaf).
//
// MODE 0  out = epi(sum_k w_k x[col_k])                 w from the packed entries (GCNConv / SimpleConv)
// MODE 1  out = epi(sum_k alpha[b,k] x[col_k])          per-sample weights, plan order (GATConv forward)
// MODE 2  dz  = sum_k alpha_t[b,k] dout[i_k] + (sum_k g_t[b,k]) att_src + da_dst att_dst;  da_src = sum_k g_t
//                                                        (GATConv backward, sender-grouped plan)
// MODE 3  dalpha_k = <dout_i, z[col_k]>; g_k = alr_k (dalpha_k - sum_k alpha_k dalpha_k); da_dst = sum_k g_k;
//         g_t[b, f2t[k]] = g_k                           (GATConv backward, receiver-grouped plan)
#pragma once
#include <algorithm>

#include "tile.cuh"

namespace gcl {

#ifndef GCL_WS_STAGES
#define GCL_WS_STAGES 3
#endif
#ifndef GCL_WS_CW
#define GCL_WS_CW 16
#endif
#ifndef GCL_WS_PW
#define GCL_WS_PW 12
#endif
constexpr int kWsStages = GCL_WS_STAGES;
constexpr int kWsConsumerWarps = GCL_WS_CW;
constexpr int kWsProducerWarps = GCL_WS_PW;
constexpr int kWsThreads = 32 * (kWsConsumerWarps + kWsProducerWarps);
constexpr int kWsProducerThreads = 32 * kWsProducerWarps;

struct WsSmem {
  int xs, ds, ent, wa, wb, dal, f2t, re, rid, desc, stage, us, rs, bars, total;   // byte offsets / sizes
};
inline WsSmem ws_smem(const TileArgs& p, int C, int SB, int mode, int elem_bytes = 4) {
  WsSmem s;
  const int E = (p.max_entries + 3) & ~3, R = p.max_rows, U = p.max_union;
  int o = 0;
  s.xs = o;   o += SB * (U + 1) * C * elem_bytes;
  s.ds = o;   o += mode == 3 ? SB * R * C * elem_bytes : 0;
  s.ent = o;  o += E * 8;
  s.wa = o;   o += mode >= 1 ? SB * E * 4 : 0;
  s.wb = o;   o += mode >= 2 ? SB * E * 4 : 0;
  s.dal = o;  o += mode == 3 ? SB * E * 4 : 0;
  s.f2t = o;  o += mode == 3 ? E * 4 : 0;
  s.re = o;   o += ((R + 1 + 3) & ~3) * 4;
  s.rid = o;  o += ((R + 3) & ~3) * 4;
  s.desc = o; o += 32;
  s.stage = (o + 127) & ~127;
  s.us = kWsStages * s.stage;                              // source-row ids of the next two items
  s.rs = s.us + 2 * ((U + 3) & ~3) * 4;                    // MODE 3: row ids of the next two items
  s.bars = s.rs + (mode == 3 ? 2 * ((R + 3) & ~3) * 4 : 0);
  s.total = s.bars + 16 * kWsStages;
  return s;
}

struct WsParams {
  TileArgs p;
  const int2* ent;                 // [n_entries] {lidx (pads / masked: max_union), weight bits}
  const float* x;                  // rows gathered through the union: x (0), z (1, 3), dout (2)
  int64_t x_bstride;
  const float* x2;                 // MODE 3: dout, rows of the tile's own rows
  const float* wa;                 // per-sample plan-order scalars: alpha (1), alpha_t (2), alpha_f (3)
  const float* wb;                 //                                 g_t (2), alr_f (3)
  int64_t w_bstride;
  float* out;                      // out (0, 1), dz (2)
  int64_t out_bstride;
  float* z_out;
  const float* bias;
  const float* prelu_slope;
  const float* att_src;            // MODE 2
  const float* att_dst;
  const float* da_dst_in;
  float* da_src_out;
  float* g_t;                      // MODE 3
  int64_t g_bstride;
  const int32_t* f2t;
  float* da_dst_out;
  int64_t n_nodes;
  int C, B, n_items, dbg;
};

// acc.{x,y} += w * v.{x,y}; acc.{z,w} += w * v.{z,w}   (two FFMA2)
__device__ __forceinline__ void fma4_packed(float4& acc, float w, const float4& v) {
  unsigned long long a0, a1, v0, v1, ww;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ww) : "f"(w), "f"(w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(acc.x), "f"(acc.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(acc.z), "f"(acc.w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(v0) : "f"(v.x), "f"(v.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(v1) : "f"(v.z), "f"(v.w));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(v0), "l"(ww));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(v1), "l"(ww));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a0));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(a1));
}
__device__ __forceinline__ void ws_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ws_cp8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ws_cp4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src) : "memory");
}
// the executing thread arrives on `bar` once all its earlier cp.async copies have landed (the arrival is part of
// the barrier's expected count: .noinc)
__device__ __forceinline__ void ws_cp_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ws_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
// 8 bf16 in a 16-byte word -> channels 0..3 / 4..7 as fp32 (a bf16 is the top half of an fp32)
__device__ __forceinline__ float4 bf16x4_lo(const uint4& u) {
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ float4 bf16x4_hi(const uint4& u) {
  return make_float4(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u), __uint_as_float(u.w << 16),
                     __uint_as_float(u.w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint4 pack_bf16x8(const float4& a, const float4& b) {
  return make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}
__device__ __forceinline__ float4 ws_prelu4(float4 a, float s) {
  return make_float4(prelu_f(a.x, s), prelu_f(a.y, s), prelu_f(a.z, s), prelu_f(a.w, s));
}

// LC = lanes that cover a row in 16-byte chunks (copy mapping); the reductions use L = LC/2 lanes per row, each
// owning words gl and gl + L (LC = 4: one word per lane)
template <int LC, int SB, int MODE, bool BF16 = false>
__global__ void __launch_bounds__(kWsThreads, 1) ws_kernel(const __grid_constant__ WsParams q, const WsSmem sm) {
  // BF16 (MODE 0 only): feature rows are bf16 (x, out, z_out point to __nv_bfloat16 data, strides in elements);
  // a 16-byte word then holds 8 channels, accumulation stays fp32
  extern __shared__ __align__(128) unsigned char smem[];
  const TileArgs& p = q.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = q.C, B = q.B, n_items = q.n_items;
  constexpr int ES = BF16 ? 2 : 4;                 // bytes per element
  const int words = (C * ES) >> 4;                 // 16-byte words per row
  const int T = p.n_tiles;
  const int zrow = p.max_union;                    // index of the all-zero row of a stage
  const int EP = (p.max_entries + 3) & ~3;         // entry capacity of the per-sample scalar planes
  const int n_my = blockIdx.x < n_items ? (n_items - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const uint32_t rowb = (uint32_t)C * (uint32_t)ES;
  const uint32_t sstride = (uint32_t)(zrow + 1) * rowb;
  const uint32_t dstride = (uint32_t)p.max_rows * rowb;
  const uint32_t smem_base = tile_smem_u32(smem);
  const uint32_t full0 = smem_base + sm.bars, empty0 = full0 + 8 * kWsStages;

  for (int idx = tid; idx < kWsStages * SB * words; idx += kWsThreads) {      // the all-zero rows
    const int st = idx / (SB * words), r = idx - st * SB * words, s = r / words, wd = r - s * words;
    *reinterpret_cast<float4*>(smem + st * sm.stage + sm.xs + (uint32_t)s * sstride + (uint32_t)zrow * rowb + wd * 16) =
        make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid == 0) {
    for (int st = 0; st < kWsStages; ++st) {
      tile_mbar_init(full0 + 8 * st, kWsProducerThreads);    // every producer lane's copies arrive
      tile_mbar_init(empty0 + 8 * st, kWsConsumerWarps);     // one arrival per consumer warp
    }
  }
  __syncthreads();

  if (warp < kWsProducerWarps) {
    // ------------------------------------------------------------------------------------------- producers
    // The producer warps share an item: copy group g = ptid / LC takes rows g, g + NG, ...; a warp requests exactly
    // the source ids its own groups will use, so its private cp.async groups + __syncwarp order them.
    constexpr int NG = kWsProducerThreads / LC;              // rows per pass of all producer threads
    constexpr int RW = 32 / LC;                              // rows per warp-wide copy instruction
    const int ptid = tid;
    const int part = lane & (LC - 1), g = ptid / LC;
    const bool clive = part < words;
    const int usz = ((zrow + 3) & ~3) * 4, rsz = ((p.max_rows + 3) & ~3) * 4;
    auto load_desc = [&](int j, int4& a, int4& b) {
      if (j < n_my) {
        const int4* dp = reinterpret_cast<const int4*>(p.tile_desc) + 2 * ((blockIdx.x + j * (int)gridDim.x) % T);
        a = __ldg(dp);                                       // {r0, nr, u0, nu}
        b = __ldg(dp + 1);                                   // {e0, ne, 0, 0}
      } else {
        a = b = make_int4(0, 0, 0, 0);
      }
    };
    auto request_ids = [&](int j, const int4& a) {           // this warp's source-row (and row) ids of item j
      for (int qi = lane;; qi += 32) {
        const int u = (qi / RW) * NG + warp * RW + (qi % RW);
        if (u >= a.w) break;
        ws_cp4(smem_base + sm.us + (j & 1) * usz + 4 * u, p.usrc + a.z + u);
      }
      if (MODE == 3) {
        for (int qi = lane;; qi += 32) {
          const int r = (qi / RW) * NG + warp * RW + (qi % RW);
          if (r >= a.y) break;
          ws_cp4(smem_base + sm.rs + (j & 1) * rsz + 4 * r, p.rows + a.x + r);
        }
      }
      cp_async_commit();
    };
    int4 da, db, na, nbq, fa, fb;                            // descriptors of items j, j+1, j+2
    load_desc(0, da, db);
    load_desc(1, na, nbq);
    request_ids(0, da);
    for (int j = 0; j < n_my; ++j) {
      const int slot = j % kWsStages;
      request_ids(j + 1, na);                                // commit order: ids(j+1) before rows(j)
      load_desc(j + 2, fa, fb);                              // lands while this item's copies are issued
      const int item = blockIdx.x + j * gridDim.x;
      const int b0 = (item / T) * SB;
      const int nb = min(SB, B - b0);
      tile_mbar_wait(empty0 + 8 * slot, (uint32_t)(((j / kWsStages) & 1) ^ 1));   // consumers released the stage
      if (j == 0) cp_async_wait<1>();              // ids(j) landed (ids(j+1) and rows(j-1) may still be in flight)
      else cp_async_wait<2>();
      __syncwarp();
      const uint32_t st = smem_base + (uint32_t)slot * (uint32_t)sm.stage;
      if (clive && !(q.dbg & 1)) {
        // per-sample source pointers once per item; a row then costs one 32-bit multiply and, per sample, one
        // 64-bit add and the copy (n_nodes * C < 2^31 is checked by the host)
        const int32_t* us = reinterpret_cast<const int32_t*>(smem + sm.us + (j & 1) * usz);
        const unsigned char* xp[SB];                         // byte pointers: fp32 or bf16 rows
#pragma unroll
        for (int s = 0; s < SB; ++s)
          xp[s] = reinterpret_cast<const unsigned char*>(q.x) + ((int64_t)(b0 + min(s, nb - 1)) * q.x_bstride) * ES + part * 16;
        const uint32_t dst0 = st + sm.xs + part * 16;
        if (nb == SB) {
#pragma unroll 4
          for (int u = g; u < da.w; u += NG) {
            const uint32_t off = (uint32_t)us[u] * rowb;
            const uint32_t dst = dst0 + (uint32_t)u * rowb;
#pragma unroll
            for (int s = 0; s < SB; ++s) ws_cp16(dst + s * sstride, xp[s] + off);
          }
        } else {
          for (int u = g; u < da.w; u += NG) {
            const uint32_t off = (uint32_t)us[u] * rowb;
            const uint32_t dst = dst0 + (uint32_t)u * rowb;
#pragma unroll
            for (int s = 0; s < SB; ++s)
              if (s < nb) ws_cp16(dst + s * sstride, xp[s] + off);
          }
        }
        if (MODE == 3) {                                     // the tile's own rows of the second tensor (dout)
          const int32_t* rs = reinterpret_cast<const int32_t*>(smem + sm.rs + (j & 1) * rsz);
          const uint32_t dd0 = st + sm.ds + part * 16;
          for (int r = g; r < da.y; r += NG) {
            const uint32_t off = (uint32_t)rs[r] * (uint32_t)C;
#pragma unroll
            for (int s = 0; s < SB; ++s)
              if (s < nb)
                ws_cp16(dd0 + (uint32_t)r * rowb + s * dstride, q.x2 + (int64_t)(b0 + s) * q.x_bstride + part * 4 + off);
          }
        }
      }
      for (int i = ptid; i <= da.y; i += kWsProducerThreads) {
        ws_cp4(st + sm.re + 4 * i, p.eptr + da.x + i);
        if (i < da.y) ws_cp4(st + sm.rid + 4 * i, p.rows + da.x + i);
      }
      for (int e = 2 * ptid; e < db.y; e += 2 * kWsProducerThreads) {      // entry pairs
        ws_cp16(st + sm.ent + 8 * e, q.ent + db.x + e);
        if (MODE >= 1) {
#pragma unroll
          for (int s = 0; s < SB; ++s) {
            if (s < nb) {
              const int64_t wo = (int64_t)(b0 + s) * q.w_bstride + db.x + e;
              ws_cp8(st + sm.wa + 4 * (s * EP + e), q.wa + wo);
              if (MODE >= 2) ws_cp8(st + sm.wb + 4 * (s * EP + e), q.wb + wo);
            }
          }
        }
        if (MODE == 3) ws_cp8(st + sm.f2t + 4 * e, q.f2t + db.x + e);
      }
      if (ptid < 2)
        ws_cp16(st + sm.desc + 16 * ptid,
                reinterpret_cast<const int4*>(p.tile_desc) + 2 * ((blockIdx.x + j * (int)gridDim.x) % T) + ptid);
      ws_cp_arrive(full0 + 8 * slot);
      cp_async_commit();
      da = na; db = nbq; na = fa; nbq = fb;
    }
    cp_async_wait<0>();
    return;
  }

  // --------------------------------------------------------------------------------------------- consumers
  constexpr int L = LC >= 8 ? LC / 2 : LC;
  constexpr int WPL = LC >= 8 ? 2 : 1;
  constexpr int kGroups = kWsConsumerWarps * 32 / L;
  const int ctid = tid - kWsProducerThreads;
  const int gl = ctid & (L - 1), grp = ctid / L;
  const bool live0 = gl < words, live1 = WPL == 2 && gl + L < words;
  const uint32_t woff0 = live0 ? gl * 16 : 0, woff1 = live1 ? (gl + L) * 16 : 0;
  float4 bv0 = make_float4(0.f, 0.f, 0.f, 0.f), bv1 = bv0;         // MODE 0/1: bias;  MODE 2: att_src
  float4 cv0 = bv0, cv1 = bv0;                                      // MODE 2: att_dst
  if (MODE <= 1 && q.bias && !BF16) {
    if (live0) bv0 = ldg4(q.bias + gl * 4);
    if (live1) bv1 = ldg4(q.bias + (gl + L) * 4);
  }
  if (MODE == 2) {
    if (live0) { bv0 = ldg4(q.att_src + gl * 4); cv0 = ldg4(q.att_dst + gl * 4); }
    if (live1) { bv1 = ldg4(q.att_src + (gl + L) * 4); cv1 = ldg4(q.att_dst + (gl + L) * 4); }
  }
  const float slope = (MODE <= 1 && q.prelu_slope) ? __ldg(q.prelu_slope) : 0.f;
  const unsigned gmask = tile_group_mask<L>(lane);

  for (int i = 0; i < n_my; ++i) {
    const int slot = i % kWsStages;
    const int item = blockIdx.x + i * gridDim.x;
    const int b0 = (item / T) * SB;
    const int nb = min(SB, B - b0);
    tile_mbar_wait(full0 + 8 * slot, (uint32_t)((i / kWsStages) & 1));
    unsigned char* st = smem + slot * sm.stage;
    const int nr = (q.dbg & 2) ? 0 : *reinterpret_cast<const int*>(st + sm.desc + 4);
    const int e0 = *reinterpret_cast<const int*>(st + sm.desc + 16);
    const int32_t* re = reinterpret_cast<const int32_t*>(st + sm.re);
    const int32_t* rid = reinterpret_cast<const int32_t*>(st + sm.rid);
    const int2* en = reinterpret_cast<const int2*>(st + sm.ent);
    const float* was = reinterpret_cast<const float*>(st + sm.wa);
    const float* wbs = reinterpret_cast<const float*>(st + sm.wb);
    const unsigned char* xb[SB];     // ragged last sample block: sample nb-1 stands in (its result is not stored)
#pragma unroll
    for (int s = 0; s < SB; ++s) xb[s] = st + sm.xs + (uint32_t)min(s, nb - 1) * sstride;

    if (BF16) {
      // bf16 rows: a word = 8 channels; fp32 accumulators, bf16 results (round to nearest even)
      for (int r = grp; r < nr; r += kGroups) {
        const int le0 = re[r] - e0, le1 = re[r + 1] - e0;
        const int64_t row = rid[r];
        float4 acc[SB][WPL][2];
#pragma unroll
        for (int s = 0; s < SB; ++s)
#pragma unroll
          for (int wq = 0; wq < WPL; ++wq) acc[s][wq][0] = acc[s][wq][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int le = le0; le < le1; le += 2) {
          const int4 e2 = *reinterpret_cast<const int4*>(en + le);
          const uint32_t o0 = (uint32_t)e2.x * rowb, o1 = (uint32_t)e2.z * rowb;
          const float w0 = __int_as_float(e2.y), w1 = __int_as_float(e2.w);
#pragma unroll
          for (int s = 0; s < SB; ++s) {
#pragma unroll
            for (int wq = 0; wq < WPL; ++wq) {
              const uint32_t wo = wq ? woff1 : woff0;
              const uint4 ua = *reinterpret_cast<const uint4*>(xb[s] + o0 + wo);
              const uint4 ub = *reinterpret_cast<const uint4*>(xb[s] + o1 + wo);
              fma4_packed(acc[s][wq][0], w0, bf16x4_lo(ua));
              fma4_packed(acc[s][wq][1], w0, bf16x4_hi(ua));
              fma4_packed(acc[s][wq][0], w1, bf16x4_lo(ub));
              fma4_packed(acc[s][wq][1], w1, bf16x4_hi(ub));
            }
          }
        }
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          if (s >= nb) break;
          const int64_t o = ((int64_t)(b0 + s) * q.out_bstride + row * C) * 2;     // byte offset
#pragma unroll
          for (int wq = 0; wq < WPL; ++wq) {
            const int wd = wq ? gl + L : gl;
            if (wq ? live1 : live0) {
              float4 a0 = acc[s][wq][0], a1 = acc[s][wq][1];
              if (q.bias) {
                const float4 b0v = ldg4(q.bias + wd * 8), b1v = ldg4(q.bias + wd * 8 + 4);
                a0 = make_float4(a0.x + b0v.x, a0.y + b0v.y, a0.z + b0v.z, a0.w + b0v.w);
                a1 = make_float4(a1.x + b1v.x, a1.y + b1v.y, a1.z + b1v.z, a1.w + b1v.w);
              }
              if (q.z_out) *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(q.z_out) + o + wd * 16) = pack_bf16x8(a0, a1);
              if (q.prelu_slope) {
                a0 = ws_prelu4(a0, slope);
                a1 = ws_prelu4(a1, slope);
              }
              *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(q.out) + o + wd * 16) = pack_bf16x8(a0, a1);
            }
          }
        }
      }
    } else if (MODE <= 2) {
      for (int r = grp; r < nr; r += kGroups) {
        const int le0 = re[r] - e0, le1 = re[r + 1] - e0;             // even count: rows are padded to pairs
        const int64_t row = rid[r];
        float dad[SB];
        if (MODE == 2) {
#pragma unroll
          for (int s = 0; s < SB; ++s) dad[s] = __ldg(q.da_dst_in + (int64_t)(b0 + min(s, nb - 1)) * q.n_nodes + row);
        }
        float4 acc0[SB], acc1[SB];
        float gsv[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          acc0[s] = acc1[s] = make_float4(0.f, 0.f, 0.f, 0.f);
          gsv[s] = 0.f;
        }
#pragma unroll(SB * WPL >= 4 ? 1 : 2)
        for (int le = le0; le < le1; le += 2) {
          const int4 e2 = *reinterpret_cast<const int4*>(en + le);    // two entries {lidx (pads: zero row), w}
          const uint32_t o0 = (uint32_t)e2.x * rowb, o1 = (uint32_t)e2.z * rowb;
          float w0 = __int_as_float(e2.y), w1 = __int_as_float(e2.w);
#pragma unroll
          for (int s = 0; s < SB; ++s) {
            if (MODE >= 1) {
              const float2 wv = *reinterpret_cast<const float2*>(was + min(s, nb - 1) * EP + le);
              w0 = wv.x; w1 = wv.y;
              if (MODE == 2) {
                const float2 gv = *reinterpret_cast<const float2*>(wbs + min(s, nb - 1) * EP + le);
                gsv[s] += gv.x;
                gsv[s] += gv.y;
              }
            }
            const float4 va = *reinterpret_cast<const float4*>(xb[s] + o0 + woff0);
            const float4 vb = *reinterpret_cast<const float4*>(xb[s] + o1 + woff0);
            fma4_packed(acc0[s], w0, va);
            fma4_packed(acc0[s], w1, vb);
            if (WPL == 2) {
              const float4 vc = *reinterpret_cast<const float4*>(xb[s] + o0 + woff1);
              const float4 vd = *reinterpret_cast<const float4*>(xb[s] + o1 + woff1);
              fma4_packed(acc1[s], w0, vc);
              fma4_packed(acc1[s], w1, vd);
            }
          }
        }
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          if (s >= nb) break;
          if (MODE <= 1) {
            const int64_t o = (int64_t)(b0 + s) * q.out_bstride + row * C;
            if (live0) {
              float4 a = make_float4(acc0[s].x + bv0.x, acc0[s].y + bv0.y, acc0[s].z + bv0.z, acc0[s].w + bv0.w);
              if (q.z_out) st4(q.z_out + o + gl * 4, a);
              if (q.prelu_slope) a = ws_prelu4(a, slope);
              st4(q.out + o + gl * 4, a);
            }
            if (live1) {
              float4 a = make_float4(acc1[s].x + bv1.x, acc1[s].y + bv1.y, acc1[s].z + bv1.z, acc1[s].w + bv1.w);
              if (q.z_out) st4(q.z_out + o + (gl + L) * 4, a);
              if (q.prelu_slope) a = ws_prelu4(a, slope);
              st4(q.out + o + (gl + L) * 4, a);
            }
          } else {
            const int64_t node = (int64_t)(b0 + s) * q.n_nodes + row;
            if (gl == 0) q.da_src_out[node] = gsv[s];
            if (live0) {
              fma4(acc0[s], gsv[s], bv0);
              fma4(acc0[s], dad[s], cv0);
              st4(q.out + node * C + gl * 4, acc0[s]);
            }
            if (live1) {
              fma4(acc1[s], gsv[s], bv1);
              fma4(acc1[s], dad[s], cv1);
              st4(q.out + node * C + (gl + L) * 4, acc1[s]);
            }
          }
        }
      }
    } else {
      // MODE 3: edge dot products by a butterfly transpose-reduce over batches of NB neighbours, softmax backward
      constexpr int NB = L / SB;
      float* dal = reinterpret_cast<float*>(st + sm.dal);
      const int32_t* f2t = reinterpret_cast<const int32_t*>(st + sm.f2t);
      const unsigned char* db_[SB];
#pragma unroll
      for (int s = 0; s < SB; ++s) db_[s] = st + sm.ds + (uint32_t)min(s, nb - 1) * dstride;
      const int jj = gl / SB, s_me = gl % SB;        // after the transpose-reduce this lane owns (neighbour jj, sample s_me)
      const bool s_ok = s_me < nb;
      const int sc = min(s_me, nb - 1);
      for (int r = grp; r < nr; r += kGroups) {
        const int le0 = re[r] - e0, le1 = re[r + 1] - e0;
        const int64_t row = rid[r];
        float4 dv0[SB], dv1[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          dv0[s] = live0 ? *reinterpret_cast<const float4*>(db_[s] + (uint32_t)r * rowb + woff0) : make_float4(0.f, 0.f, 0.f, 0.f);
          dv1[s] = live1 ? *reinterpret_cast<const float4*>(db_[s] + (uint32_t)r * rowb + woff1) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float tp = 0.f;
        for (int base = le0; base < le1; base += NB) {
          float pr[L];
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const int le = min(base + j, le1 - 1);   // past the row end: re-read the last entry, dropped below
            const uint32_t o = (uint32_t)en[le].x * rowb;
#pragma unroll
            for (int s = 0; s < SB; ++s) {
              float d = dot4(dv0[s], *reinterpret_cast<const float4*>(xb[s] + o + woff0));
              if (WPL == 2) d += dot4(dv1[s], *reinterpret_cast<const float4*>(xb[s] + o + woff1));
              pr[j * SB + s] = d;
            }
          }
          const float rsum = tile_xreduce<L>(pr, gl, gmask);
          const int le = base + jj;
          if (le < le1 && s_ok) {
            tp += was[sc * EP + le] * rsum;
            dal[sc * EP + le] = rsum;
          }
        }
        const float t = tile_gsum_strided<L, SB>(tp, gmask);
        float gs = 0.f;
        for (int base = le0; base < le1; base += NB) {
          const int le = base + jj;
          if (le < le1 && s_ok) {
            const float gk = wbs[sc * EP + le] * (dal[sc * EP + le] - t);
            const int et = f2t[le];
            if (et >= 0) q.g_t[(int64_t)(b0 + s_me) * q.g_bstride + et] = gk;
            gs += gk;
          }
        }
        gs = tile_gsum_strided<L, SB>(gs, gmask);
        if (jj == 0 && s_ok) q.da_dst_out[(int64_t)(b0 + s_me) * q.n_nodes + row] = gs;
      }
    }
    __syncwarp();
    if (lane == 0) ws_arrive(empty0 + 8 * slot);                     // this warp is done with the stage
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
inline int ws_pick_sb(const TileArgs& p, int C, int B, int mode) {
  static const int sb_cap = getenv("GCL_TILE_SB") ? atoi(getenv("GCL_TILE_SB")) : 4;
  int sb = mode == 3 ? std::min(sb_cap, 2) : sb_cap;
  while (sb > 1 && (sb > B || ws_smem(p, C, sb, mode).total > 227 * 1024)) sb >>= 1;
  return ws_smem(p, C, sb, mode).total > 227 * 1024 ? 0 : sb;
}

template <int LC, int SB, int MODE>
int ws_launch(WsParams q, cudaStream_t s, const char* what) {
  const WsSmem sm = ws_smem(q.p, q.C, SB, MODE);
  auto kern = ws_kernel<LC, SB, MODE>;
  static bool attr_set = false;               // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, what);
    attr_set = true;
  }
  const int64_t n_items = (int64_t)q.p.n_tiles * ceil_div(q.B, SB);
  static const int ctas = getenv("GCL_TILE_CTAS") ? atoi(getenv("GCL_TILE_CTAS")) : kNumSMs;
  static const int dbg = getenv("GCL_TILE_DBG") ? atoi(getenv("GCL_TILE_DBG")) : 0;   // elimination runs (wrong results)
  q.n_items = (int)n_items;
  q.dbg = dbg;
  const unsigned grid = (unsigned)std::min<int64_t>(n_items, ctas);
  kern<<<grid, kWsThreads, sm.total, s>>>(q, sm);
  GCL_CHECK_LAUNCH(what);
  return GCL_OK;
}

template <int LC, int MODE>
int ws_dispatch_l(const WsParams& q, cudaStream_t s, const char* what) {
  const int sb = ws_pick_sb(q.p, q.C, q.B, MODE);
  if (sb == 0) {
    set_error("%s: a tile (union %d rows of %d channels) does not fit the shared-memory ring", what, q.p.max_union, q.C);
    return GCL_ERR_UNSUPPORTED;
  }
  if (MODE != 3 && sb == 4) return ws_launch<LC, MODE == 3 ? 2 : 4, MODE>(q, s, what);
  if (sb >= 2) return ws_launch<LC, 2, MODE>(q, s, what);
  return ws_launch<LC, 1, MODE>(q, s, what);
}

// bf16 feature rows (MODE 0): words are 8 channels wide
template <int LC, int SB>
int ws_launch_bf16(WsParams q, cudaStream_t s, const char* what) {
  const WsSmem sm = ws_smem(q.p, q.C, SB, 0, 2);
  auto kern = ws_kernel<LC, SB, 0, true>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, what);
    attr_set = true;
  }
  const int64_t n_items = (int64_t)q.p.n_tiles * ceil_div(q.B, SB);
  q.n_items = (int)n_items;
  q.dbg = 0;
  kern<<<(unsigned)std::min<int64_t>(n_items, kNumSMs), kWsThreads, sm.total, s>>>(q, sm);
  GCL_CHECK_LAUNCH(what);
  return GCL_OK;
}

template <int LC>
int ws_dispatch_bf16_l(const WsParams& q, cudaStream_t s, const char* what) {
  int sb = 2;
  while (sb > 1 && (sb > q.B || ws_smem(q.p, q.C, sb, 0, 2).total > 227 * 1024)) sb >>= 1;
  if (ws_smem(q.p, q.C, sb, 0, 2).total > 227 * 1024) {
    set_error("%s: a tile (union %d rows of %d bf16 channels) does not fit the shared-memory ring", what, q.p.max_union, q.C);
    return GCL_ERR_UNSUPPORTED;
  }
  if (sb == 2) return ws_launch_bf16<LC, 2>(q, s, what);
  return ws_launch_bf16<LC, 1>(q, s, what);
}

inline int ws_dispatch_bf16(const WsParams& q, cudaStream_t s, const char* what) {
  const int words = q.C / 8;
  if (words <= 4) return ws_dispatch_bf16_l<4>(q, s, what);
  if (words <= 8) return ws_dispatch_bf16_l<8>(q, s, what);
  if (words <= 16) return ws_dispatch_bf16_l<16>(q, s, what);
  return ws_dispatch_bf16_l<32>(q, s, what);
}

template <int MODE>
int ws_dispatch(const WsParams& q, cudaStream_t s, const char* what) {
  const int words = q.C / 4;
  if (words <= 4) return ws_dispatch_l<4, MODE>(q, s, what);
  if (words <= 8) return ws_dispatch_l<8, MODE>(q, s, what);
  if (words <= 16) return ws_dispatch_l<16, MODE>(q, s, what);
  return ws_dispatch_l<32, MODE>(q, s, what);
}

}  // namespace gcl

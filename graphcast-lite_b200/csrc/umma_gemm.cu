// K7 on the 5th-generation tensor cores: fp32-accurate dense transforms with tcgen05 (UMMA) in
// 3xTF32 -- every fp32 operand is split into hi = tf32(x) and lo = tf32(x - hi) and
//     D += A_hi B_hi + A_lo B_hi + A_hi B_lo           (fp32 accumulation in TMEM)
// which keeps ~22 mantissa bits per product, enough for the rel 1e-4 parity bar that plain TF32 (10
// bits) misses (DESIGN.md "dense transform").  The split is why operands are staged by ordinary loads +
// st.shared instead of TMA: the loader warps convert while they copy, writing the UMMA canonical
// 128-byte-swizzled layouts directly.
//
//   umma_linear_kernel : C[M,N] = A[M,K] W[N,K]^T (+bias, PReLU, pre-activation copy)
//                        forward of nn.Linear / PyG lin (W = weight) and dX = dY W (W = weight^T)
//   umma_dw_kernel     : P[s][N1,N2] = sum_{r in slice s} A[r,N1]^T B[r,N2]   (dW = dY^T X partials)
//
// Persistent, warp-specialised CTAs (one per SM): warps 0-3 epilogue (TMEM -> registers -> global; a
// warp may only touch TMEM lanes 32*(warp%4)..+31), warp 4 allocates TMEM and its lane 0 issues the
// MMAs, warps 5-11 load/split/stage.  mbarrier rings: full/empty per smem stage, full/empty per TMEM
// accumulator (two accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1).
#include "common.cuh"

namespace gcl {
namespace {

constexpr int kTileM = 128;
constexpr int kKB = 32;                        // fp32 elements per 128-byte swizzle row
constexpr int kPartBytes = kTileM * 128;       // one 128-row x 128-byte operand block (hi or lo): 16 KB
// 12 warps = 3 per SM sub-partition, which lets every thread have up to 168 registers (13 warps would put
// 4 on one sub-partition and cap everything at 128: the loaders' prefetch ring then spills).
constexpr int kEpiWarps = 4, kLoadWarps = 7;
constexpr int kThreads = (kEpiWarps + 1 + kLoadWarps) * 32;   // 384
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kMaxSmem = 227 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets TMEM lane (base lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ bool al16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// one contiguous, 16 B aligned chunk smem -> global through the bulk-copy (TMA) engine
__device__ __forceinline__ void bulk_store(float* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// hi = x rounded to tf32's 10 mantissa bits (add half an ulp to the bit pattern, then mask: round to
// nearest, ties away, in two integer ops), lo = x - hi exactly (one FADD, either sign).  The tensor core
// ignores lo's low 13 bits, so hi + lo carries ~21-22 mantissa bits of x and the dropped lo*lo term has no
// preferred sign.  (Plain truncation is one op cheaper but biases every product low by ~1e-6; cvt.rna.tf32
// gives the same values at ~4x the instructions, and the loaders are issue-bound.)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}
__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x);
  split_tf32(v.y, h.y, l.y);
  split_tf32(v.z, h.z, l.z);
  split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = l;
}

// UMMA shared-memory descriptor, 128-byte swizzle, version 1 (sm_100): start address, LBO, SBO in 16 B
// units (cute::UMMA::SmemDescriptor bit layout).
// layout: 2 = SWIZZLE_128B (K-major operands), 1 = SWIZZLE_128B_BASE32B (the only layout tcgen05 accepts for
// MN-major tf32 operands: 32-byte chunks of a 128-byte row XORed with row % 4, atoms of 4 rows).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint64_t layout = 2) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (layout << 61);
}
// cute::UMMA::InstrDescriptor: D = F32, A = B = TF32, M = 128, N; major bits 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a [rows x 128 B] swizzled block
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return (uint32_t)(r * 128 + (((c ^ (r & 7)) & 7) << 4)); }

// ------------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] W[N,K]^T  (K-major operands).  VEC: K % 4 == 0 and 16 B aligned A / W rows.
template <bool VEC>
__global__ void __launch_bounds__(kThreads, 1)
    umma_linear_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C, int64_t M,
                       int N, int K, int n_pad, int nkb, int nst, int tmem_cols, int cw, int stage_z,
                       const float* __restrict__ bias, const float* __restrict__ prelu_slope,
                       float* __restrict__ z_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_block = n_pad * 128;                       // one K-block of W (hi or lo)
  uint8_t* b_hi = smem;
  uint8_t* b_lo = b_hi + (size_t)nkb * b_block;
  uint8_t* a_st = b_lo + (size_t)nkb * b_block;
  float* bias_s = reinterpret_cast<float*>(a_st + (size_t)nst * 2 * kPartBytes);       // n_pad floats
  const int pitch = cw * 4 + 16;                         // staging row pitch: +16 B keeps st.shared.v4 conflict-free
  uint8_t* c_stage = reinterpret_cast<uint8_t*>(bias_s + n_pad);                        // [128][pitch]
  uint8_t* z_stage = c_stage + (size_t)(cw ? kTileM * pitch : 0);                      // [128][pitch] if stage_z
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(z_stage + (size_t)(stage_z ? kTileM * pitch : 0)) + 15) & ~uintptr_t(15));
  const uint32_t bars = smem_u32(bar_ptr);
  // barrier indices: full[s] = s, empty[s] = nst + s, tmem_full[a] = 2 nst + a, tmem_empty[a] = 2 nst + 2 + a
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + 2 * nst + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + kTileM - 1) / kTileM;

  if (tid == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(bars + 8 * s, kLoadThreads);
      mbar_init(bars + 8 * (nst + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bars + 8 * (2 * nst + a), 1);
      mbar_init(bars + 8 * (2 * nst + 2 + a), kEpiWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps) tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols);

  // weights -> hi/lo, swizzled, resident for the whole kernel
  for (int idx = tid; idx < nkb * n_pad * 8; idx += kThreads) {
    const int c = idx & 7, n = (idx >> 3) % n_pad, kb = (idx >> 3) / n_pad;
    const int k = kb * kKB + c * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
      const float* src = W + (int64_t)n * K + k;
      if (VEC) {
        if (k < K) v = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        if (k + 0 < K) v.x = __ldg(src + 0);
        if (k + 1 < K) v.y = __ldg(src + 1);
        if (k + 2 < K) v.z = __ldg(src + 2);
        if (k + 3 < K) v.w = __ldg(src + 3);
      }
    }
    split_store(b_hi + (size_t)kb * b_block, b_lo + (size_t)kb * b_block, sw_off(n, c), v);
  }
  for (int n = tid; n < n_pad; n += kThreads) bias_s[n] = (bias && n < N) ? __ldg(bias + n) : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp > kEpiWarps) {
    // ===================== loaders: global A -> registers -> hi/lo -> swizzled smem stage =====================
    const int lt = tid - (kEpiWarps + 1) * 32;           // 0..223
    const int64_t nitems = ntiles * nkb;
    int64_t it_local = 0;
    constexpr int kPer = (1024 + kLoadThreads - 1) / kLoadThreads;   // 16-byte chunks per thread per stage (5)
    auto fetch = [&](int64_t item, float4 (&dst)[kPer]) {
      const int64_t tile = blockIdx.x + (item / nkb) * gridDim.x;
      const int kb = (int)(item % nkb);
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const int id = lt + i * kLoadThreads;              // 0..1023 valid
        const int r = id >> 3, c = id & 7;
        const int64_t row = tile * kTileM + r;
        const int k = kb * kKB + c * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id < 1024 && row < M) {
          const float* src = A + row * K + k;
          if (VEC) {
            if (k < K) v = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            if (k + 0 < K) v.x = __ldg(src + 0);
            if (k + 1 < K) v.y = __ldg(src + 1);
            if (k + 2 < K) v.z = __ldg(src + 2);
            if (k + 3 < K) v.w = __ldg(src + 3);
          }
        }
        dst[i] = v;
      }
    };
    // number of (tile, kblock) items of this CTA
    const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t my_items = my_tiles * nkb;
    (void)nitems;
    // kDepth register buffers form a ring: a buffer is refilled right after it has been staged, so
    // kDepth items (kDepth x 16 KB per SM) are always in flight -- that, not the MMA, sets the pace.
    constexpr int kDepth = 5;
    float4 buf[kDepth][kPer];
#pragma unroll
    for (int d = 0; d < kDepth; ++d)
      if (d < my_items) fetch(d, buf[d]);
    for (int64_t base = 0; base < my_items; base += kDepth) {
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        it_local = base + d;
        if (it_local < my_items) {
          const int st = (int)(it_local % nst);
          const uint32_t ph = (uint32_t)((it_local / nst) & 1);
          mbar_wait(bars + 8 * (nst + st), ph ^ 1);
          uint8_t* hi = a_st + (size_t)st * 2 * kPartBytes;
          uint8_t* lo = hi + kPartBytes;
#pragma unroll
          for (int i = 0; i < kPer; ++i) {
            const int id = lt + i * kLoadThreads;
            if (id < 1024) split_store(hi, lo, sw_off(id >> 3, id & 7), buf[d][i]);
          }
          fence_proxy_async();
          mbar_arrive(bars + 8 * st);
          if (it_local + kDepth < my_items) fetch(it_local + kDepth, buf[d]);
        }
      }
    }
  } else if (warp == kEpiWarps) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(n_pad, 0, 0);
      int64_t it_local = 0, t_local = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_local) {
        const int acc = (int)(t_local & 1);
        mbar_wait(bars + 8 * (2 * nst + 2 + acc), (uint32_t)(((t_local >> 1) & 1) ^ 1));   // accumulator drained
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * n_pad);
        for (int kb = 0; kb < nkb; ++kb, ++it_local) {
          const int st = (int)(it_local % nst);
          mbar_wait(bars + 8 * st, (uint32_t)((it_local / nst) & 1));
          tc_fence_after();
          const uint32_t a_hi = smem_u32(a_st + (size_t)st * 2 * kPartBytes);
          const uint32_t a_lo = a_hi + kPartBytes;
          const uint32_t bh = smem_u32(b_hi + (size_t)kb * b_block);
          const uint32_t bl = smem_u32(b_lo + (size_t)kb * b_block);
          const int ksteps = min(4, (K - kb * kKB + 7) / 8);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da_hi = make_desc(a_hi + ks * 32, 16, 1024), da_lo = make_desc(a_lo + ks * 32, 16, 1024);
            const uint64_t db_hi = make_desc(bh + ks * 32, 16, 1024), db_lo = make_desc(bl + ks * 32, 16, 1024);
            umma_tf32(d_tmem, da_lo, db_hi, idesc, (kb | ks) ? 1u : 0u);   // small terms first
            umma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
            umma_tf32(d_tmem, da_hi, db_hi, idesc, 1u);
          }
          umma_commit(bars + 8 * (nst + st));                // smem stage free once these MMAs retire
        }
        umma_commit(bars + 8 * (2 * nst + acc));             // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> (+bias, PReLU) -> global =====================
    // Every thread owns one output row.  With cw > 0 the row segment is staged in this thread's own smem
    // row and leaves as ONE asynchronous bulk copy (cp.async.bulk, 16 B aligned): the direct alternative,
    // 32 lanes storing 16 B each to 32 different rows, costs 32 LSU sector transactions per instruction
    // and was the single largest term of the kernel time.
    const float slope = prelu_slope ? __ldg(prelu_slope) : 0.f;
    const int my = warp * 32 + lane;
    uint8_t* crow_s = c_stage + (size_t)my * pitch;
    uint8_t* zrow_s = z_stage + (size_t)my * pitch;
    int64_t t_local = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_local) {
      const int acc = (int)(t_local & 1);
      mbar_wait(bars + 8 * (2 * nst + acc), (uint32_t)((t_local >> 1) & 1));
      tc_fence_after();
      const int64_t row = tile * kTileM + my;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * n_pad);
      if (cw > 0) {
        for (int p = 0; p < N; p += cw) {
          bulk_wait_read();                                  // my previous copies no longer read my smem rows
          for (int c0 = p; c0 < p + cw; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (c0 + 4 * q < p + cw) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * q);
                float4 o = make_float4(v[4 * q] + b4.x, v[4 * q + 1] + b4.y, v[4 * q + 2] + b4.z, v[4 * q + 3] + b4.w);
                if (z_out) {
                  if (stage_z) *reinterpret_cast<float4*>(zrow_s + (c0 - p + 4 * q) * 4) = o;
                  else if (row < M) *reinterpret_cast<float4*>(z_out + row * N + c0 + 4 * q) = o;
                }
                if (prelu_slope)
                  o = make_float4(prelu_f(o.x, slope), prelu_f(o.y, slope), prelu_f(o.z, slope), prelu_f(o.w, slope));
                *reinterpret_cast<float4*>(crow_s + (c0 - p + 4 * q) * 4) = o;
              }
            }
          }
          fence_proxy_async();
          if (row < M) {
            bulk_store(C + row * N + p, smem_u32(crow_s), (uint32_t)cw * 4);
            if (z_out && stage_z) bulk_store(z_out + row * N + p, smem_u32(zrow_s), (uint32_t)cw * 4);
          }
          bulk_commit();
        }
      } else {
        const bool vec_ok = (N & 3) == 0 && al16_dev(C) && (!z_out || al16_dev(z_out));
        for (int c0 = 0; c0 < n_pad; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          if (row < M) {
            if (vec_ok && c0 + 16 <= N) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * q);
                float4 o = make_float4(v[4 * q] + b4.x, v[4 * q + 1] + b4.y, v[4 * q + 2] + b4.z, v[4 * q + 3] + b4.w);
                if (z_out) *reinterpret_cast<float4*>(z_out + row * N + c0 + 4 * q) = o;
                if (prelu_slope)
                  o = make_float4(prelu_f(o.x, slope), prelu_f(o.y, slope), prelu_f(o.z, slope), prelu_f(o.w, slope));
                *reinterpret_cast<float4*>(C + row * N + c0 + 4 * q) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (c0 + j < N) {
                  const float x = v[j] + bias_s[c0 + j];
                  if (z_out) z_out[row * N + c0 + j] = x;
                  C[row * N + c0 + j] = prelu_slope ? prelu_f(x, slope) : x;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bars + 8 * (2 * nst + 2 + acc));
    }
    bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// P[s][m][n] = sum over the CTA's row slice of A[r][m] * B[r][n]   (dW = dY^T X, split over rows).
// The reduction index r is the slow one in memory, i.e. both operands arrive "MN-major".  tcgen05 can
// read MN-major tf32 (SWIZZLE_128B_BASE32B), but that path measured ~3.4x below the K-major MMA rate,
// so the loaders transpose instead: a lane owns one column m and reads it for 4 consecutive rows
// (4 warp-coalesced 128-byte loads), which is exactly one 16-byte K-chunk of smem row m in the K-major
// SWIZZLE_128B layout -- consecutive lanes hit different swizzled chunks, so the stores are conflict-free.
// B gets a constant row n = N of ones, so accumulator column N is the column sum of A (bias gradient).
template <int NAU, int NBU>   // units (32 columns x 4 rows) per loader warp for A and B
__global__ void __launch_bounds__(kThreads, 1)
    umma_dw_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ part,
                   float* __restrict__ part_colsum, int64_t R, int M, int N, int n_pad, int nst, int tmem_cols,
                   int64_t rows_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_part = n_pad * 128;                          // hi or lo of the B stage: [n_pad rows x 128 B]
  const int stage_bytes = 2 * kPartBytes + 2 * b_part;     // A_hi, A_lo, B_hi, B_lo
  uint8_t* st_base = smem;
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(st_base + (size_t)nst * stage_bytes);
  const uint32_t bars = smem_u32(bar_ptr);                 // full[s] = s, empty[s] = nst + s, acc_full = 2 nst
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + 2 * nst + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t r_beg = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = min(R, r_beg + rows_per_cta);
  const int64_t nkb = r_end > r_beg ? (r_end - r_beg + kKB - 1) / kKB : 0;

  if (tid == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(bars + 8 * s, kLoadThreads);
      mbar_init(bars + 8 * (nst + s), 1);
    }
    mbar_init(bars + 8 * (2 * nst), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps) tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols);
  // constant parts of every stage: zero everything, then the row of ones (hi = 1, lo = 0) at n = N
  for (int i = tid; i < nst * stage_bytes / 16; i += kThreads)
    reinterpret_cast<float4*>(st_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int i = tid; i < nst * 8; i += kThreads) {
    uint8_t* b_hi = st_base + (size_t)(i >> 3) * stage_bytes + 2 * kPartBytes;
    *reinterpret_cast<float4*>(b_hi + sw_off(N, i & 7)) = make_float4(1.f, 1.f, 1.f, 1.f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp > kEpiWarps) {
    const int lw = warp - (kEpiWarps + 1);                 // 0..6
    const int na_units = 8 * ((M + 31) / 32), nb_units = 8 * ((N + 31) / 32);
    // register prefetch ring: as deep as the 168-register budget allows for this shape
    constexpr int kDepth = (NAU + NBU <= 6) ? 4 : ((NAU + NBU <= 8) ? 3 : 2);
    float4 abuf[kDepth][NAU], bbuf[kDepth][NBU];
    // per-unit element offsets relative to the first row of a K-block (-1 = unit or column out of range)
    int offa[NAU], offb[NBU];
#pragma unroll
    for (int i = 0; i < NAU; ++i) {
      const int u = lw + kLoadWarps * i, m = (u >> 3) * 32 + lane;
      offa[i] = (u < na_units && m < M) ? (u & 7) * 4 * M + m : -1;
    }
#pragma unroll
    for (int i = 0; i < NBU; ++i) {
      const int u = lw + kLoadWarps * i, n = (u >> 3) * 32 + lane;
      offb[i] = (u < nb_units && n < N) ? (u & 7) * 4 * N + n : -1;
    }
    auto fetch = [&](int64_t kb, float4 (&da)[NAU], float4 (&db)[NBU]) {
      const int64_t r0 = r_beg + kb * kKB;
      const float* pa = A + r0 * M;
      const float* pb = B + r0 * N;
      if (r0 + kKB <= r_end) {                              // full K-block: no row guards
#pragma unroll
        for (int i = 0; i < NAU; ++i) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (offa[i] >= 0) {
            const float* src = pa + offa[i];
            v = make_float4(__ldg(src), __ldg(src + M), __ldg(src + 2 * M), __ldg(src + 3 * M));
          }
          da[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NBU; ++i) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (offb[i] >= 0) {
            const float* src = pb + offb[i];
            v = make_float4(__ldg(src), __ldg(src + N), __ldg(src + 2 * N), __ldg(src + 3 * N));
          }
          db[i] = v;
        }
      } else {                                              // last, partial K-block of the slice
        const int left = (int)(r_end - r0);
#pragma unroll
        for (int i = 0; i < NAU; ++i) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (offa[i] >= 0) {
            const int rr = ((lw + kLoadWarps * i) & 7) * 4;
            const float* src = pa + offa[i];
            if (rr + 0 < left) v.x = __ldg(src);
            if (rr + 1 < left) v.y = __ldg(src + M);
            if (rr + 2 < left) v.z = __ldg(src + 2 * M);
            if (rr + 3 < left) v.w = __ldg(src + 3 * M);
          }
          da[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NBU; ++i) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (offb[i] >= 0) {
            const int rr = ((lw + kLoadWarps * i) & 7) * 4;
            const float* src = pb + offb[i];
            if (rr + 0 < left) v.x = __ldg(src);
            if (rr + 1 < left) v.y = __ldg(src + N);
            if (rr + 2 < left) v.z = __ldg(src + 2 * N);
            if (rr + 3 < left) v.w = __ldg(src + 3 * N);
          }
          db[i] = v;
        }
      }
    };
#pragma unroll
    for (int d = 0; d < kDepth; ++d)
      if (d < nkb) fetch(d, abuf[d], bbuf[d]);
    for (int64_t base = 0; base < nkb; base += kDepth) {
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        const int64_t kb = base + d;
        if (kb < nkb) {
          const int st = (int)(kb % nst);
          mbar_wait(bars + 8 * (nst + st), (uint32_t)(((kb / nst) & 1) ^ 1));
          uint8_t* a_hi = st_base + (size_t)st * stage_bytes;
          uint8_t* a_lo = a_hi + kPartBytes;
          uint8_t* b_hi = a_lo + kPartBytes;
          uint8_t* b_lo = b_hi + b_part;
#pragma unroll
          for (int i = 0; i < NAU; ++i) {
            const int u = lw + kLoadWarps * i;
            if (offa[i] >= 0) split_store(a_hi, a_lo, sw_off((u >> 3) * 32 + lane, u & 7), abuf[d][i]);
          }
#pragma unroll
          for (int i = 0; i < NBU; ++i) {
            const int u = lw + kLoadWarps * i;
            if (offb[i] >= 0) split_store(b_hi, b_lo, sw_off((u >> 3) * 32 + lane, u & 7), bbuf[d][i]);
          }
          fence_proxy_async();
          mbar_arrive(bars + 8 * st);
          if (kb + kDepth < nkb) fetch(kb + kDepth, abuf[d], bbuf[d]);
        }
      }
    }
  } else if (warp == kEpiWarps) {
    if (lane == 0 && nkb > 0) {
      const uint32_t idesc = make_idesc(n_pad, 0, 0);
      for (int64_t kb = 0; kb < nkb; ++kb) {
        const int st = (int)(kb % nst);
        mbar_wait(bars + 8 * st, (uint32_t)((kb / nst) & 1));
        tc_fence_after();
        const uint32_t a_hi = smem_u32(st_base + (size_t)st * stage_bytes);
        const uint32_t a_lo = a_hi + kPartBytes;
        const uint32_t b_hi = a_lo + kPartBytes;
        const uint32_t b_lo = b_hi + b_part;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t da_hi = make_desc(a_hi + ks * 32, 16, 1024), da_lo = make_desc(a_lo + ks * 32, 16, 1024);
          const uint64_t db_hi = make_desc(b_hi + ks * 32, 16, 1024), db_lo = make_desc(b_lo + ks * 32, 16, 1024);
          umma_tf32(tmem_base, da_lo, db_hi, idesc, (kb | ks) ? 1u : 0u);
          umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
          umma_tf32(tmem_base, da_hi, db_hi, idesc, 1u);
        }
        umma_commit(bars + 8 * (nst + st));
      }
      umma_commit(bars + 8 * (2 * nst));
    }
  } else {
    const int m = warp * 32 + lane;
    float* prow = part + ((int64_t)blockIdx.x * M + m) * N;
    if (nkb > 0) {
      mbar_wait(bars + 8 * (2 * nst), 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < n_pad; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (m < M) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = c0 + j;
            if (n < N) prow[n] = v[j];
            else if (n == N && part_colsum) part_colsum[(int64_t)blockIdx.x * M + m] = v[j];
          }
        }
      }
    } else if (m < M) {
      for (int n = 0; n < N; ++n) prow[n] = 0.f;
      if (part_colsum) part_colsum[(int64_t)blockIdx.x * M + m] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

struct LinPlan {
  int n_pad, nkb, nst, tmem_cols, cw, stage_z;
  size_t smem;
  bool ok;
};

LinPlan plan_linear(int64_t N, int64_t K, bool has_z, bool c_aligned) {
  LinPlan p{};
  p.n_pad = (int)((N + 15) / 16 * 16);
  p.nkb = (int)((K + kKB - 1) / kKB);
  p.ok = N >= 1 && p.n_pad <= 256 && K >= 1;
  if (!p.ok) return p;
  const long b_bytes = 2L * p.nkb * p.n_pad * 128;
  const long fixed = 1024 /*align slack*/ + 256 /*barriers*/ + 16 + 4L * p.n_pad /*bias*/;
  int cands[8], nc = 0;
  if (N % 4 == 0 && c_aligned) {
    cands[nc++] = (int)N;   // one bulk copy per row and tile; partial-row passes would serialise on wait_group.read
  }
  cands[nc++] = 0;   // direct stores
  int best_nst = 0;
  for (int want = 3; want >= 2 && best_nst == 0; --want) {
    for (int i = 0; i < nc && best_nst == 0; ++i) {
      for (int sz = (has_z && cands[i] > 0) ? 1 : 0; sz >= 0 && best_nst == 0; --sz) {
        const long staging = cands[i] ? (1L + sz) * kTileM * (cands[i] * 4 + 16) : 0;
        const long room = (long)kMaxSmem - b_bytes - fixed - staging;
        int nst = (int)(room / (2 * kPartBytes));
        if (nst > 6) nst = 6;
        if (nst >= want) {
          best_nst = nst;
          p.cw = cands[i];
          p.stage_z = sz;
          p.smem = (size_t)(b_bytes + fixed + staging + (long)nst * 2 * kPartBytes);
        }
      }
    }
  }
  p.nst = best_nst;
  p.ok = best_nst >= 2;
  int cols = 32;
  while (cols < 2 * p.n_pad) cols <<= 1;
  p.tmem_cols = cols;
  return p;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace


// Returns GCL_OK when launched, GCL_ERR_UNSUPPORTED when the shape does not fit (caller falls back to
// the FFMA kernel), or an error.
int umma_linear(const float* A, const float* W_nk, float* C, int64_t M, int64_t N, int64_t K, const float* bias,
                const float* slope, float* z_out, cudaStream_t s) {
  LinPlan p = plan_linear(N, K, z_out != nullptr, al16(C) && (!z_out || al16(z_out)));
  if (!p.ok || M <= 0) return GCL_ERR_UNSUPPORTED;
  const bool vec = (K % 4 == 0) && al16(A) && al16(W_nk);
  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  cudaError_t e;
  if (vec) {
    e = cudaFuncSetAttribute(umma_linear_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return fail_cuda(e, "umma_linear(smem attr)");
    umma_linear_kernel<true><<<grid, kThreads, p.smem, s>>>(A, W_nk, C, M, (int)N, (int)K, p.n_pad, p.nkb, p.nst,
                                                            p.tmem_cols, p.cw, p.stage_z, bias, slope, z_out);
  } else {
    e = cudaFuncSetAttribute(umma_linear_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return fail_cuda(e, "umma_linear(smem attr)");
    umma_linear_kernel<false><<<grid, kThreads, p.smem, s>>>(A, W_nk, C, M, (int)N, (int)K, p.n_pad, p.nkb, p.nst,
                                                             p.tmem_cols, p.cw, p.stage_z, bias, slope, z_out);
  }
  GCL_CHECK_LAUNCH("umma_linear");
  return GCL_OK;
}

namespace {
struct DwUmmaPlan {
  int n_pad, nst, tmem_cols, grid;
  int64_t rows_per_cta;
  size_t smem;
  bool ok;
};
DwUmmaPlan plan_dw(int64_t R, int64_t M, int64_t N) {
  DwUmmaPlan p{};
  p.ok = R > 0 && M >= 1 && M <= kTileM && N >= 1 && N <= 128;
  if (!p.ok) return p;
  p.n_pad = (int)((N + 1 + 15) / 16 * 16);
  const size_t stage = 2 * (size_t)kPartBytes + 2 * (size_t)p.n_pad * 128;
  const size_t fixed = 1024 + 256;
  p.nst = (int)(((size_t)kMaxSmem - fixed) / stage);
  if (p.nst > 4) p.nst = 4;
  p.ok = p.nst >= 2;
  p.smem = (size_t)p.nst * stage + fixed;
  int cols = 32;
  while (cols < p.n_pad) cols <<= 1;
  p.tmem_cols = cols;
  int64_t grid = (R + kKB - 1) / kKB;
  if (grid > kNumSMs) grid = kNumSMs;
  p.rows_per_cta = ((R + grid - 1) / grid + kKB - 1) / kKB * kKB;
  p.grid = (int)((R + p.rows_per_cta - 1) / p.rows_per_cta);
  return p;
}
}  // namespace

// number of row slices (= partial tiles) umma_dw will write, 0 if the shape is unsupported
int umma_dw_splits(int64_t R, int64_t M, int64_t N) {
  DwUmmaPlan p = plan_dw(R, M, N);
  return p.ok ? p.grid : 0;
}

// part[s][M][N], part_colsum[s][M] (nullable) for s < umma_dw_splits(); A [R,M], B [R,N]
int umma_dw(const float* A, const float* B, float* part, float* part_colsum, int64_t R, int64_t M, int64_t N,
            cudaStream_t s) {
  DwUmmaPlan p = plan_dw(R, M, N);
  if (!p.ok) return GCL_ERR_UNSUPPORTED;
  auto go = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return fail_cuda(e, "umma_dw(smem attr)");
    kern<<<p.grid, kThreads, p.smem, s>>>(A, B, part, part_colsum, R, (int)M, (int)N, p.n_pad, p.nst, p.tmem_cols,
                                          p.rows_per_cta);
    return GCL_OK;
  };
  // (32-column x 4-row) units per loader warp: 8 * ceil(width / 32) units over 7 warps
  const int au = (8 * (int)((M + 31) / 32) + kLoadWarps - 1) / kLoadWarps;
  const int bu = (8 * (int)((N + 31) / 32) + kLoadWarps - 1) / kLoadWarps;
  int rc;
  if (au <= 3 && bu <= 3) rc = go(umma_dw_kernel<3, 3>);
  else if (au <= 3) rc = go(umma_dw_kernel<3, 5>);
  else if (bu <= 3) rc = go(umma_dw_kernel<5, 3>);
  else rc = go(umma_dw_kernel<5, 5>);
  if (rc != GCL_OK) return rc;
  GCL_CHECK_LAUNCH("umma_dw");
  return GCL_OK;
}

}  // namespace gcl

// K7 on the 5th-generation tensor cores: fp32-accurate dense transforms with tcgen05 (UMMA) in
// 3xTF32 -- every fp32 operand is split into hi (tf32) and lo = x - hi and
//     D = A_hi B_hi + (A_lo B_hi + A_hi B_lo)           (fp32 accumulation in TMEM)
// which keeps ~22 mantissa bits per product, enough for the rel 1e-4 parity bar that plain TF32 (10
// bits) misses (DESIGN.md "Dense transform").
//
//   umma_linear_tma_kernel : C[M,N] = A[M,K] W[N,K]^T (+bias, PReLU, pre-activation copy; or the PReLU-backward
//                            epilogue for dX; or the attention-score epilogue of GATConv's lin).  TMA tensor
//                            loads feed the hi operand directly (hi = trunc), converter warps write lo, the
//                            corrections accumulate in their own TMEM accumulator, TMA tensor stores.  Default.
//   umma_dw_tma_kernel     : P[s][N1,N2] = sum_{r in slice s} A[r,N1]^T B[r,N2] (dW = dY^T X partials): TMA loads
//                            raw row blocks, converter warps transpose + split them into K-major stages.  Default.
//   umma_linear_kernel, umma_dw_kernel : the register-staged first generation (loader warps ld.global -> split ->
//                            st.shared).  Kept for rows that are not 16-byte aligned (TMA cannot address them)
//                            and as the A/B switch GCL_UMMA_NO_TMA=1.
//
// All are persistent, warp-specialised CTAs (one per SM): a warp may only touch TMEM lanes 32*(warp%4)..+31, so
// epilogue warps come in groups of four; one warp allocates TMEM and one elected lane issues the MMAs; mbarrier
// rings per smem stage and per TMEM accumulator (two, so the epilogue of tile i overlaps the MMAs of tile i+1).
#include <cuda.h>   // CUtensorMap types only; the encoder is resolved at run time

#include "common.cuh"

namespace gcl {
// debugging / A-B switch (GCL_UMMA_NO_TMA=1): keep the register-staged kernels even where TMA applies
static const bool g_force_register_staging = [] {
  const char* e = getenv("GCL_UMMA_NO_TMA");
  return e && e[0] == '1';
}();
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

// two 16-column loads in flight, one wait; the wait takes the destination registers as in/out operands so no
// consumer can be scheduled above it
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr_a, uint32_t taddr_b, float (&a)[16], float (&b)[16]) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr_a));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(taddr_b));
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]),
                 "+r"(q[8]), "+r"(q[9]), "+r"(q[10]), "+r"(q[11]), "+r"(q[12]), "+r"(q[13]), "+r"(q[14]), "+r"(q[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(q[i]);
  }
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

// UMMA shared-memory descriptors (cute::UMMA::SmemDescriptor bit layout, version 1 = sm_100): start address,
// LBO, SBO in 16-byte units, layout 2 = SWIZZLE_128B for K-major operands -- built by desc_lo / desc_k128 below.
// The MMA-issuing warp is a single instruction stream, and ncu showed it -- not DRAM, not the tensor core --
// pacing the kernels when every MMA rebuilt its descriptors and ring indices with divisions.  So: the K-major
// SWIZZLE_128B descriptor is split into a constant high word and a low word (address >> 4 | LBO) that advances
// by 2 per 8-element K step, ring positions are counters, and the loop runs warp-uniform with one elected lane.
constexpr uint32_t kDescHiK128 = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));   // SBO=1024, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t desc_k128(uint32_t lo) { return ((uint64_t)kDescHiK128 << 32) | lo; }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
// the three 3xTF32 products of one 8-element K step (small terms first)
__device__ __forceinline__ void umma_3x(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                        uint32_t idesc, uint32_t accumulate) {
  umma_tf32(d_tmem, desc_k128(a_lo), desc_k128(b_hi), idesc, accumulate);
  umma_tf32(d_tmem, desc_k128(a_hi), desc_k128(b_lo), idesc, 1u);
  umma_tf32(d_tmem, desc_k128(a_hi), desc_k128(b_hi), idesc, 1u);
}

// Accumulator split used by the TMA kernels: the two correction products go to their own accumulator columns.
// The tensor core adds into its fp32 accumulator with truncation (measured: positive operands come out ~5e-7 low
// after the 24 accumulation steps of a K = 64 row); keeping the ~2^-11-sized corrections out of the main
// accumulator leaves it 8 steps, and the epilogue adds the parts with an ordinary rounded fp32 add.  With the hi
// and lo copies of the N-side operand adjacent in shared memory, A_hi x [B_hi | B_lo] is a single MMA of width
// 2 n_pad (main | correction columns adjacent in TMEM), followed by A_lo x B_hi into the correction columns.

// cute::UMMA::InstrDescriptor: D = F32, A = B = TF32, M = 128, N; major bits 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major, int m = kTileM) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t idesc = make_idesc(n_pad, 0, 0);
    const uint32_t a_lo0 = desc_lo(smem_u32(a_st)), bh_lo = desc_lo(smem_u32(b_hi)), bl_lo = desc_lo(smem_u32(b_lo));
    const uint32_t b_step = (uint32_t)b_block >> 4;
    int st = 0;
    uint32_t ph = 0;
    int64_t t_local = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_local) {
      const int acc = (int)(t_local & 1);
      mbar_wait(bars + 8 * (2 * nst + 2 + acc), (uint32_t)(((t_local >> 1) & 1) ^ 1));   // accumulator drained
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * n_pad);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(bars + 8 * st, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ah = a_lo0 + (uint32_t)st * (2 * kPartBytes >> 4), al = ah + (kPartBytes >> 4);
          const uint32_t bh = bh_lo + (uint32_t)kb * b_step, bl = bl_lo + (uint32_t)kb * b_step;
          const int ksteps = min(4, (K - kb * kKB + 7) / 8);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (ks < ksteps) umma_3x(d_tmem, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
          umma_commit(bars + 8 * (nst + st));                // smem stage free once these MMAs retire
          if (kb == nkb - 1) umma_commit(bars + 8 * (2 * nst + acc));   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++st == nst) { st = 0; ph ^= 1; }
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
    const uint32_t idesc = make_idesc(n_pad, 0, 0);
    const uint32_t st_lo = desc_lo(smem_u32(st_base));
    int st = 0;
    uint32_t ph = 0;
    for (int64_t kb = 0; kb < nkb; ++kb) {
      mbar_wait(bars + 8 * st, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ah = st_lo + (uint32_t)st * ((uint32_t)stage_bytes >> 4), al = ah + (kPartBytes >> 4);
        const uint32_t bh = al + (kPartBytes >> 4), bl = bh + ((uint32_t)b_part >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_3x(tmem_base, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
        umma_commit(bars + 8 * (nst + st));
        if (kb == nkb - 1) umma_commit(bars + 8 * (2 * nst));
      }
      __syncwarp();
      if (++st == nst) { st = 0; ph ^= 1; }
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

// ------------------------------------------------------------------------------------------------
// TMA-fed variant of the linear kernel (the default whenever rows are 16-byte aligned).
//
// The register-staged kernel above moves every byte through the LSU twice and leaves the tile as 128
// separate 256-byte bulk copies (UBLKCP is issued per lane, serially): ncu showed neither DRAM nor the tensor
// core busy, just long-scoreboard and bulk-copy issue stalls at ~3.4 TB/s.  Here the copy engine does the
// moving, in 16 KB boxes:
//   * TMA loads [128 rows x 32 floats] slabs of A straight into the K-major SWIZZLE_128B layout.  The raw
//     fp32 slab IS the hi operand: kind::tf32 reads the top 19 bits of each word, i.e. hi = trunc_tf32(x);
//   * six converter warps compute lo = x - trunc_tf32(x) (exact) smem -> smem at the same swizzled offset,
//     one LDS.128 + one STS.128 per four elements, conflict-free;
//   * the epilogue writes the output tile into SWIZZLE_128B slabs (conflict-free by construction) and one
//     thread sends it with one TMA tensor store per 32-column slab; row / column tails are clipped by the
//     tensor map, so there is no tail code at all.
// hi by truncation biases the dropped lo(A) lo(W) term (<= 2^-21 relative per product, see split_tf32);
// W keeps the rounded split.
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float lo_trunc(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

constexpr int kSlabBytes = kTileM * 128;     // [128 rows x 32 fp32], 16 KB
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32x4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// roles (14 warps): 0-7 epilogue (warp w owns TMEM lanes 32 (w % 4).. and the 16-column blocks with
// block % 2 == w / 4), 8 MMA issue + TMEM allocation, 9 TMA producer, 10-13 converters
constexpr int kLtEpiWarps = 8, kLtConvWarps = 4;
constexpr int kLtMmaWarp = kLtEpiWarps, kLtTmaWarp = kLtEpiWarps + 1;
constexpr int kLtThreads = (kLtEpiWarps + 2 + kLtConvWarps) * 32;      // 448
constexpr int kLtConvThreads = kLtConvWarps * 32, kLtEpiThreads = kLtEpiWarps * 32;

__global__ void __launch_bounds__(kLtThreads, 1)
    umma_linear_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmC,
                           const __grid_constant__ CUtensorMap tmZ, const float* __restrict__ W, int64_t M, int N,
                           int K, int n_pad, int nkb, int nob, int s_raw, int s_lo, int nbuf, int tmem_cols, int has_z,
                           const float* __restrict__ bias, const float* __restrict__ prelu_slope, int n_split,
                           const __grid_constant__ CUtensorMap tmZin, const float* __restrict__ act_slope,
                           float* __restrict__ dslope_part, const float* __restrict__ att,
                           float* __restrict__ sc_src, float* __restrict__ sc_dst) {
  // att != nullptr (the `lin` of a single-head GATConv): the epilogue also produces the attention logits' node
  // terms sc_src[row] = <y[row], att[0:N]>, sc_dst[row] = <y[row], att[N:2N]> (models.py:336-357 via PyG's
  // (x * att).sum(-1)), which otherwise costs a second pass over y.
  // act_slope != nullptr (used for dX): the result is multiplied by PReLU'(z_in) of the layer that produced this
  // layer's input, z_in tiles arrive by TMA like A, and sum(result * min(z_in, 0)) -- that PReLU's slope
  // gradient -- leaves as one partial per CTA.  Saves the separate pass over dX, z_in and dZ.
  const int has_act = act_slope != nullptr;
  // n_split = 2: W (hi + lo) of a wide layer does not fit next to the operand ring, so CTA 2c and 2c + 1 walk the
  // same row tiles and each owns one half of the output columns (N is then the half width).  The second read of
  // a tile comes out of L2: HBM traffic stays what it was.
  const int half = blockIdx.x % n_split, cta = blockIdx.x / n_split, ncta = gridDim.x / n_split;
  W += (int64_t)half * N * K;
  if (bias) bias += half * N;
  const int col0 = half * N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_block = n_pad * 128;
  // W per K block: [hi rows | lo rows] adjacent, so that A_hi x [W_hi | W_lo] is one MMA of width 2 n_pad whose
  // left half lands in the main and whose right half in the correction accumulator (adjacent TMEM columns)
  uint8_t* b_hi = smem;
  uint8_t* b_lo = b_hi + b_block;
  uint8_t* raw = b_hi + (size_t)nkb * 2 * b_block;
  uint8_t* lo = raw + (size_t)s_raw * kSlabBytes;
  uint8_t* c_out = lo + (size_t)s_lo * kSlabBytes;                     // [nbuf][1 + has_z][nob] slabs
  const int out_buf_bytes = (1 + has_z) * nob * kSlabBytes;
  uint8_t* zin = c_out + (size_t)nbuf * out_buf_bytes;                 // [nob] slabs of z_in, single-buffered
  float* bias_s = reinterpret_cast<float*>(zin + (size_t)(has_act ? nob : 0) * kSlabBytes);
  float* att_s = bias_s + n_pad;                                       // [2][n_pad]
  float* sc_part = att_s + 2 * n_pad;                                  // [2 tiles][2 column groups][128 rows][2]
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sc_part + 2 * 2 * kTileM * 2) + 15) & ~uintptr_t(15));
  const uint32_t bars = smem_u32(bar_ptr);
  const int kRawFull = 0, kRawEmpty = s_raw, kLoFull = 2 * s_raw, kLoEmpty = 2 * s_raw + s_lo,
            kAccFull = 2 * s_raw + 2 * s_lo, kAccEmpty = kAccFull + 2;
  const int kZinFull = kAccEmpty + 2, kZinEmpty = kZinFull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + kZinEmpty + 1);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int64_t my_tiles = (ntiles > cta) ? (ntiles - cta + ncta - 1) / ncta : 0;
  const int64_t my_items = my_tiles * nkb;

  if (tid == 0) {
    for (int s = 0; s < s_raw; ++s) {
      mbar_init(bar(kRawFull + s), 1);
      mbar_init(bar(kRawEmpty + s), 1);
    }
    for (int s = 0; s < s_lo; ++s) {
      mbar_init(bar(kLoFull + s), kLtConvThreads);
      mbar_init(bar(kLoEmpty + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar(kAccFull + a), 1);
      mbar_init(bar(kAccEmpty + a), kLtEpiThreads);
    }
    mbar_init(bar(kZinFull), 1);
    mbar_init(bar(kZinEmpty), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kLtMmaWarp) tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols);
  for (int idx = tid; idx < nkb * n_pad * 8; idx += kLtThreads) {       // W -> rounded hi/lo, resident
    const int c = idx & 7, n = (idx >> 3) % n_pad, kb = (idx >> 3) / n_pad;
    const int k = kb * kKB + c * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N && k < K) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * K + k));
    split_store(b_hi + (size_t)kb * 2 * b_block, b_lo + (size_t)kb * 2 * b_block, sw_off(n, c), v);
  }
  for (int n = tid; n < n_pad; n += kLtThreads) bias_s[n] = (bias && n < N) ? __ldg(bias + n) : 0.f;
  for (int n = tid; n < 2 * n_pad; n += kLtThreads) {
    const int which = n / n_pad, c = n % n_pad;
    att_s[n] = (att && c < N) ? __ldg(att + which * N + c) : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kLtTmaWarp) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int64_t tile = cta; tile < ntiles; tile += ncta) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar(kRawEmpty + s), ph);
          mbar_expect_tx(bar(kRawFull + s), kSlabBytes);
          tma_load_2d(smem_u32(raw + (size_t)s * kSlabBytes), &tmA, kb * kKB, (int)(tile * kTileM), bar(kRawFull + s));
          if (++s == s_raw) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp > kLtTmaWarp) {
    // ===================== converters: lo = x - trunc_tf32(x), same swizzled offset =====================
    const int ct = tid - (kLtTmaWarp + 1) * 32;          // 0..127
    int s = 0, l = 0;
    uint32_t sph = 0, lph = 1;
    for (int64_t it = 0; it < my_items; ++it) {
      mbar_wait(bar(kRawFull + s), sph);
      mbar_wait(bar(kLoEmpty + l), lph);
      const float4* src = reinterpret_cast<const float4*>(raw + (size_t)s * kSlabBytes);
      float4* dst = reinterpret_cast<float4*>(lo + (size_t)l * kSlabBytes);
#pragma unroll
      for (int i = 0; i < 1024 / kLtConvThreads; ++i) {
        const float4 v = src[ct + i * kLtConvThreads];
        dst[ct + i * kLtConvThreads] = make_float4(lo_trunc(v.x), lo_trunc(v.y), lo_trunc(v.z), lo_trunc(v.w));
      }
      fence_proxy_async();
      mbar_arrive(bar(kLoFull + l));
      if (++s == s_raw) { s = 0; sph ^= 1; }
      if (++l == s_lo) { l = 0; lph ^= 1; }
    }
  } else if (warp == kLtMmaWarp) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t idesc = make_idesc(n_pad, 0, 0), idesc_wide = make_idesc(2 * n_pad, 0, 0);
    const uint32_t raw_lo = desc_lo(smem_u32(raw)), lo_lo = desc_lo(smem_u32(lo));
    const uint32_t bh_lo = desc_lo(smem_u32(b_hi));
    const uint32_t b_step = (uint32_t)(2 * b_block) >> 4;
    int rs = 0, ls = 0;
    uint32_t rph = 0, lph = 0;
    int64_t t_local = 0;
    for (int64_t tile = cta; tile < ntiles; tile += ncta, ++t_local) {
      const int acc = (int)(t_local & 1);
      mbar_wait(bar(kAccEmpty + acc), (uint32_t)(((t_local >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * n_pad), d_corr = d_tmem + (uint32_t)n_pad;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(bar(kRawFull + rs), rph);
        mbar_wait(bar(kLoFull + ls), lph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ah = raw_lo + (uint32_t)rs * (kSlabBytes >> 4), al = lo_lo + (uint32_t)ls * (kSlabBytes >> 4);
          const uint32_t bh = bh_lo + (uint32_t)kb * b_step;
          const int ksteps = min(4, (K - kb * kKB + 7) / 8);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < ksteps) {
              umma_tf32(d_tmem, desc_k128(ah + 2 * ks), desc_k128(bh + 2 * ks), idesc_wide, (kb | ks) ? 1u : 0u);
              umma_tf32(d_corr, desc_k128(al + 2 * ks), desc_k128(bh + 2 * ks), idesc, 1u);
            }
          }
          umma_commit(bar(kRawEmpty + rs));
          umma_commit(bar(kLoEmpty + ls));
          if (kb == nkb - 1) umma_commit(bar(kAccFull + acc));
        }
        __syncwarp();
        if (++rs == s_raw) { rs = 0; rph ^= 1; }
        if (++ls == s_lo) { ls = 0; lph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> swizzled smem slabs -> TMA tensor store ========
    // With two output buffers the only synchronisation per tile is one named barrier: thread 0 waits (before
    // it) until the PREVIOUS tile's stores have finished reading their buffer, which is the one the NEXT tile
    // writes.  With one buffer (not enough smem for two) that wait has to sit in front of the writes.
    const float slope = prelu_slope ? __ldg(prelu_slope) : 0.f;
    const float aslope = has_act ? __ldg(act_slope) : 0.f;
    float dsl = 0.f;
    const int my = (warp & 3) * 32 + lane, cgrp = warp >> 2;
    // z_in tiles are fetched by epilogue thread 0 itself, one tile ahead (right after the barrier that says every
    // thread is done with the previous one), so the A producer never waits on the epilogue
    auto fetch_zin = [&](int64_t tile) {
      mbar_expect_tx(bar(kZinFull), (uint32_t)nob * kSlabBytes);
      for (int j = 0; j < nob; ++j)
        tma_load_2d(smem_u32(zin + (size_t)j * kSlabBytes), &tmZin, col0 + j * kKB, (int)(tile * kTileM), bar(kZinFull));
    };
    if (has_act && tid == 0 && cta < ntiles) fetch_zin(cta);
    int64_t t_local = 0;
    for (int64_t tile = cta; tile < ntiles; tile += ncta, ++t_local) {
      const int acc = (int)(t_local & 1);
      uint8_t* ob = c_out + (size_t)(nbuf == 2 ? (t_local & 1) : 0) * out_buf_bytes;
      uint8_t* zb = ob + (size_t)nob * kSlabBytes;
      if (nbuf == 1 && t_local > 0) {
        if (tid == 0) bulk_wait_read();
        named_bar(1, kLtEpiThreads);
      }
      mbar_wait(bar(kAccFull + acc), (uint32_t)((t_local >> 1) & 1));
      tc_fence_after();
      if (has_act) mbar_wait(bar(kZinFull), (uint32_t)(t_local & 1));
      const bool row_ok = tile * kTileM + my < M;          // rows past M hold zero-filled z_in: keep them out of dsl
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(acc * 2 * n_pad);
      float ps = 0.f, pd = 0.f;
      for (int c0 = cgrp * 16; c0 < n_pad; c0 += 32) {
        float v[16], vc[16];
        tmem_ld16x2(taddr + c0, taddr + n_pad + c0, v, vc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = c0 + 4 * q;
          const uint32_t off = (uint32_t)(col >> 5) * kSlabBytes + sw_off(my, (col & 31) >> 2);
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + col);
          float4 o = make_float4((v[4 * q] + vc[4 * q]) + b4.x, (v[4 * q + 1] + vc[4 * q + 1]) + b4.y,
                                 (v[4 * q + 2] + vc[4 * q + 2]) + b4.z, (v[4 * q + 3] + vc[4 * q + 3]) + b4.w);
          if (att) {
            const float4 s4 = *reinterpret_cast<const float4*>(att_s + col), d4 = *reinterpret_cast<const float4*>(att_s + n_pad + col);
            ps += (o.x * s4.x + o.y * s4.y) + (o.z * s4.z + o.w * s4.w);
            pd += (o.x * d4.x + o.y * d4.y) + (o.z * d4.z + o.w * d4.w);
          }
          if (has_act) {
            const float4 zv = *reinterpret_cast<const float4*>(zin + off);
            if (row_ok && col < N)
              dsl += (zv.x > 0.f ? 0.f : o.x * zv.x) + (zv.y > 0.f ? 0.f : o.y * zv.y) + (zv.z > 0.f ? 0.f : o.z * zv.z) +
                     (zv.w > 0.f ? 0.f : o.w * zv.w);
            o = make_float4(zv.x > 0.f ? o.x : aslope * o.x, zv.y > 0.f ? o.y : aslope * o.y, zv.z > 0.f ? o.z : aslope * o.z,
                            zv.w > 0.f ? o.w : aslope * o.w);
          }
          if (has_z) *reinterpret_cast<float4*>(zb + off) = o;
          if (prelu_slope)
            o = make_float4(prelu_f(o.x, slope), prelu_f(o.y, slope), prelu_f(o.z, slope), prelu_f(o.w, slope));
          *reinterpret_cast<float4*>(ob + off) = o;
        }
      }
      float* scp = sc_part + (size_t)(t_local & 1) * (2 * kTileM * 2);
      if (att) {
        scp[(cgrp * kTileM + my) * 2] = ps;
        scp[(cgrp * kTileM + my) * 2 + 1] = pd;
      }
      tc_fence_before();
      mbar_arrive(bar(kAccEmpty + acc));
      fence_proxy_async();
      if (nbuf == 2 && tid == 0) bulk_wait_read();
      named_bar(1, kLtEpiThreads);
      if (att && cgrp == 0 && row_ok) {                    // column group 0 + 1, fixed order
        sc_src[tile * kTileM + my] = scp[my * 2] + scp[(kTileM + my) * 2];
        sc_dst[tile * kTileM + my] = scp[my * 2 + 1] + scp[(kTileM + my) * 2 + 1];
      }
      if (tid == 0) {
        if (has_act && tile + ncta < ntiles) fetch_zin(tile + ncta);   // everyone has read this tile's z_in
        for (int j = 0; j < nob; ++j) {
          tma_store_2d(&tmC, col0 + j * kKB, (int)(tile * kTileM), smem_u32(ob + (size_t)j * kSlabBytes));
          if (has_z) tma_store_2d(&tmZ, col0 + j * kKB, (int)(tile * kTileM), smem_u32(zb + (size_t)j * kSlabBytes));
        }
        bulk_commit();
      }
    }
    if (has_act) {                                         // per-CTA slope-gradient partial, fixed order
      __shared__ float dsl_s[kLtEpiWarps];
      dsl = warp_sum(dsl);
      if (lane == 0) dsl_s[warp] = dsl;
      named_bar(1, kLtEpiThreads);
      if (tid == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kLtEpiWarps; ++w) t += dsl_s[w];
        dslope_part[blockIdx.x] = t;
      }
    }
    if (tid == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kLtMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Weights-stationary variant for the wide layers (64 < N <= 128, K <= 128): C^T = W X^T.
//
// In the kernel above the resident W (hi + lo) of a 128 -> 128 layer takes 128 KB of shared memory, so it has to
// run column-split (n_split = 2): every A tile goes through shared memory twice (TMA write, converter read + write,
// MMA operand reads, in both CTAs of a pair) and ncu shows the shared-memory pipe, not HBM, as the bound (~1.1 MB
// of shared-memory traffic per 128 output rows at 128 B/clk).  Here the roles of the operands are swapped:
//   * W is the M-side operand and lives in TENSOR MEMORY (tcgen05.mma with the A operand in TMEM): lane = output
//     channel, 128 columns of hi and 128 of lo.  It costs no shared memory and no shared-memory bandwidth;
//   * the activations are the N-side operand: [64 rows x 32 floats] slabs, hi (raw, truncated by the tensor core)
//     and lo adjacent in one 16 KB stage, so W_hi x [X_hi | X_lo] is ONE MMA of N = 128 (main | correction
//     accumulator columns) followed by W_lo x X_hi (N = 64) into the correction columns -- the same three products;
//   * the accumulator is the transposed tile [channel lane][row column]; the epilogue thread of channel c writes
//     its 64 rows into the swizzled output slabs with 4-byte stores (32 consecutive channels of one row = 32
//     different banks) and the tile leaves by TMA tensor stores as before.
// Shared-memory traffic per 64 rows: 32 KB TMA write, 64 KB converter, 96 KB MMA operand reads, 64 KB output
// = 256 KB (1.0 us at 128 B/clk) against 1.45 us of HBM time for the tile: HBM is the bound again.
// TMEM: columns 0..127 W_hi, 128..255 W_lo, 256..383 and 384..511 two accumulators (main 64 | correction 64).
constexpr int kTsRows = 64;
constexpr int kTsSlab = kTsRows * 128;          // [64 rows x 32 fp32], 8 KB
constexpr int kTsStage = 2 * kTsSlab;           // hi | lo
constexpr int kTsWCols = 128, kTsAcc0 = 2 * kTsWCols, kTsAccCols = 2 * kTsRows;

// D[tmem] (+)= A[tmem] * B[smem]: the M-side operand is read from tensor memory (lane = row, one tf32 per column)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// thread i of the warp writes 16 consecutive columns of TMEM lane (base lane + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

// roles (17 warps): 0-7 epilogue (warp w: TMEM lanes = channels 32 (w % 4).., rows 32 (w / 4)..), 8 MMA issue +
// TMEM allocation, 9 TMA producer, 10-13 converters, 14-16 MMA issue.  Four MMA-issuing warps, tile t belongs to
// warp t % 4: with 64-row tiles a single issuing warp -- ~60 dependent uniform-datapath instructions per 16 KB
// stage -- was the critical path (ncu: that warp never idle, tensor pipe 47 %, converters waiting for data).  A
// tile's MMAs all come from one thread, so their order and the commit that follows them are unaffected.
constexpr int kTsMmaWarps = 4;
constexpr int kTsWarps = kLtEpiWarps + 2 + kLtConvWarps + (kTsMmaWarps - 1);     // 17
constexpr int kTsThreads = kTsWarps * 32;                                        // 544

template <bool HAS_Z, bool HAS_ACT>
__global__ void __launch_bounds__(kTsThreads, 1)
    umma_linear_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmC,
                          const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmZin,
                          const float* __restrict__ W, int64_t M, int N, int K, int nkb, int nob, int nst, int nbuf,
                          const float* __restrict__ bias, const float* __restrict__ prelu_slope,
                          const float* __restrict__ act_slope, float* __restrict__ dslope_part,
                          float* __restrict__ colsum_part) {
  // colsum_part (HAS_ACT only, nullable): [2 gridDim.x][N] column sums of the result over this CTA's rows, i.e.
  // the bias gradient of the layer that produced z_in -- each epilogue thread owns a channel, so it is a register
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;                                               // [nst][hi slab | lo slab]
  uint8_t* c_out = stages + (size_t)nst * kTsStage;                     // [nbuf][1 + HAS_Z][nob] slabs
  const int out_buf_bytes = (1 + (int)HAS_Z) * nob * kTsSlab;
  uint8_t* zin = c_out + (size_t)nbuf * out_buf_bytes;                  // [2][nob] slabs of z_in (two tiles ahead)
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(zin + (size_t)(HAS_ACT ? 2 * nob : 0) * kTsSlab);
  const uint32_t bars = smem_u32(bar_ptr);
  const int kRawFull = 0, kLoFull = nst, kEmpty = 2 * nst, kAccFull = 3 * nst, kAccEmpty = kAccFull + 2,
            kZinFull = kAccEmpty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + kZinFull + 2);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int64_t ntiles = (M + kTsRows - 1) / kTsRows;
  const int64_t my_tiles = (ntiles > cta) ? (ntiles - cta + ncta - 1) / ncta : 0;
  const int64_t my_items = my_tiles * nkb;

  if (tid == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(bar(kRawFull + s), 1);
      mbar_init(bar(kLoFull + s), kLtConvThreads);
      mbar_init(bar(kEmpty + s), kTsMmaWarps);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar(kAccFull + a), 1);
      mbar_init(bar(kAccEmpty + a), kLtEpiThreads);
      mbar_init(bar(kZinFull + a), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kLtMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < kLtEpiWarps) {
    // W -> tensor memory, rounded hi / exact lo; rows past N and columns past K are zero
    const int c = (warp & 3) * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int kbeg = (warp >> 2) * (kTsWCols / 2);
    for (int k0 = kbeg; k0 < kbeg + kTsWCols / 2; k0 += 16) {
      float hi[16], lo[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < N && k < K) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)c * K + k));
        split_tf32(v.x, hi[4 * q], lo[4 * q]);
        split_tf32(v.y, hi[4 * q + 1], lo[4 * q + 1]);
        split_tf32(v.z, hi[4 * q + 2], lo[4 * q + 2]);
        split_tf32(v.w, hi[4 * q + 3], lo[4 * q + 3]);
      }
      tmem_st16(tl + (uint32_t)k0, hi);
      tmem_st16(tl + (uint32_t)(kTsWCols + k0), lo);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int conv_first = kLtTmaWarp + 1, mma_extra_first = conv_first + kLtConvWarps;
  if (warp == kLtTmaWarp) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int64_t tile = cta; tile < ntiles; tile += ncta) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar(kEmpty + s), ph);
          mbar_expect_tx(bar(kRawFull + s), kTsSlab);
          tma_load_2d(smem_u32(stages + (size_t)s * kTsStage), &tmA, kb * kKB, (int)(tile * kTsRows), bar(kRawFull + s));
          if (++s == nst) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= conv_first && warp < mma_extra_first) {
    // ===================== converters: lo = x - trunc_tf32(x) into the second half of the stage ============
    const int ct = tid - conv_first * 32;          // 0..127
    int s = 0;
    uint32_t sph = 0;
    for (int64_t it = 0; it < my_items; ++it) {
      mbar_wait(bar(kRawFull + s), sph);
      const float4* src = reinterpret_cast<const float4*>(stages + (size_t)s * kTsStage);
      float4* dst = reinterpret_cast<float4*>(stages + (size_t)s * kTsStage + kTsSlab);
#pragma unroll
      for (int i = 0; i < kTsSlab / 16 / kLtConvThreads; ++i) {
        const float4 v = src[ct + i * kLtConvThreads];
        dst[ct + i * kLtConvThreads] = make_float4(lo_trunc(v.x), lo_trunc(v.y), lo_trunc(v.z), lo_trunc(v.w));
      }
      fence_proxy_async();
      mbar_arrive(bar(kLoFull + s));
      if (++s == nst) { s = 0; sph ^= 1; }
    }
  } else if (warp == kLtMmaWarp || warp >= mma_extra_first) {
    // ===================== MMA issuers: tile t_local belongs to issuer t_local % 4 =====================
    // Every issuer walks ALL items in ring order and waits on every barrier phase (an mbarrier parity wait is only
    // meaningful one phase ahead), but issues -- and commits -- only for its own tiles; for the others it just
    // signs the stage off, so a stage is released by one commit + three plain arrivals.
    const int me = (warp == kLtMmaWarp) ? 0 : warp - mma_extra_first + 1;
    const uint32_t idesc_wide = make_idesc(2 * kTsRows, 0, 0), idesc = make_idesc(kTsRows, 0, 0);
    const uint32_t st_lo = desc_lo(smem_u32(stages));
    const uint32_t w_hi = tmem_base, w_lo = tmem_base + (uint32_t)kTsWCols;
    const bool full_k = (K & (kKB - 1)) == 0;
    int s = 0;
    uint32_t ph = 0;
    for (int64_t t_local = 0; t_local < my_tiles; ++t_local) {
      const bool mine = (int)(t_local & (kTsMmaWarps - 1)) == me;
      const int acc = (int)(t_local & 1);
      mbar_wait(bar(kAccEmpty + acc), (uint32_t)(((t_local >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_main = tmem_base + (uint32_t)(kTsAcc0 + acc * kTsAccCols), d_corr = d_main + (uint32_t)kTsRows;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(bar(kRawFull + s), ph);
        mbar_wait(bar(kLoFull + s), ph);
        if (mine) {
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b = st_lo + (uint32_t)s * (kTsStage >> 4);
            const uint32_t kc = (uint32_t)(kb * kKB);
            if (full_k) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma_tf32_ts(d_main, w_hi + kc + 8 * ks, desc_k128(b + 2 * ks), idesc_wide, (kb | ks) ? 1u : 0u);
                umma_tf32_ts(d_corr, w_lo + kc + 8 * ks, desc_k128(b + 2 * ks), idesc, 1u);
              }
            } else {
              const int ksteps = min(4, (K - kb * kKB + 7) / 8);
              for (int ks = 0; ks < ksteps; ++ks) {
                umma_tf32_ts(d_main, w_hi + kc + 8 * ks, desc_k128(b + 2 * ks), idesc_wide, (kb | ks) ? 1u : 0u);
                umma_tf32_ts(d_corr, w_lo + kc + 8 * ks, desc_k128(b + 2 * ks), idesc, 1u);
              }
            }
            umma_commit(bar(kEmpty + s));
            if (kb == nkb - 1) umma_commit(bar(kAccFull + acc));
          }
          __syncwarp();
        } else if (lane == 0) {
          mbar_arrive(bar(kEmpty + s));
        }
        if (++s == nst) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM (transposed tile) -> swizzled smem slabs -> TMA tensor store =====
    const float slope = prelu_slope ? __ldg(prelu_slope) : 1.f;      // slope 1: PReLU is the identity
    const float aslope = HAS_ACT ? __ldg(act_slope) : 0.f;
    float dsl = 0.f, csum = 0.f;
    const int c = (warp & 3) * 32 + lane, rh = warp >> 2;
    const bool ch_ok = (warp & 3) < nob;                  // this warp's 32 channels have an output slab
    const float bias_c = (bias && c < N) ? __ldg(bias + c) : 0.f;
    // byte offset of (row r, channel c) in the [nob] swizzled slabs = c_off + r * 128 + swz[r & 7]
    const uint32_t c_off = (uint32_t)(c >> 5) * kTsSlab + (uint32_t)(c & 3) * 4u + (uint32_t)(rh * 32 * 128);
    uint32_t swz[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) swz[q] = c_off + (uint32_t)(((((c & 31) >> 2) ^ q) & 7) << 4);
    const uint32_t zin_s = smem_u32(zin);
    // z_in tiles arrive two tiles ahead, fetched by epilogue thread 0 right after the barrier that says every
    // thread is done with the buffer
    auto fetch_zin = [&](int64_t tile, int buf) {
      mbar_expect_tx(bar(kZinFull + buf), (uint32_t)nob * kTsSlab);
      for (int j = 0; j < nob; ++j)
        tma_load_2d(zin_s + (uint32_t)(buf * nob + j) * kTsSlab, &tmZin, j * kKB, (int)(tile * kTsRows),
                    bar(kZinFull + buf));
    };
    if (HAS_ACT && tid == 0) {
      if (cta < ntiles) fetch_zin(cta, 0);
      if (cta + ncta < ntiles) fetch_zin(cta + ncta, 1);
    }
    int64_t t_local = 0;
    for (int64_t tile = cta; tile < ntiles; tile += ncta, ++t_local) {
      const int acc = (int)(t_local & 1);
      const uint32_t ob = smem_u32(c_out + (size_t)(nbuf == 2 ? (t_local & 1) : 0) * out_buf_bytes);
      const uint32_t zb = ob + (uint32_t)nob * kTsSlab;
      const uint32_t zt = zin_s + (uint32_t)(acc * nob) * kTsSlab;
      if (nbuf == 1 && t_local > 0) {
        if (tid == 0) bulk_wait_read();
        named_bar(1, kLtEpiThreads);
      }
      mbar_wait(bar(kAccFull + acc), (uint32_t)((t_local >> 1) & 1));
      tc_fence_after();
      if (HAS_ACT) mbar_wait(bar(kZinFull + acc), (uint32_t)((t_local >> 1) & 1));
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kTsAcc0 + acc * kTsAccCols + rh * 32);
      const int64_t rows_left = M - (tile * kTsRows + rh * 32);       // rows of this half that exist
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float v[16], vc[16];
        tmem_ld16x2(taddr + 16 * i, taddr + kTsRows + 16 * i, v, vc);
        if (ch_ok) {
          float zq[16];
          if (HAS_ACT) {          // all 16 loads first: the asm loads / stores below keep their program order, and a
#pragma unroll                    // load behind every store would put a shared-memory round trip on every element
            for (int j = 0; j < 16; ++j) zq[j] = lds_f32(zt + swz[j & 7] + (uint32_t)((16 * i + j) * 128));
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t off = swz[j & 7] + (uint32_t)((16 * i + j) * 128);
            float o = (v[j] + vc[j]) + bias_c;
            if (HAS_ACT) {
              const float zv = zq[j];
              if (16 * i + j < rows_left) dsl += zv > 0.f ? 0.f : o * zv;
              o = zv > 0.f ? o : aslope * o;
              csum += o;                                  // rows past M are exact zeros (zero-filled A rows)
            }
            if (HAS_Z) sts_f32(zb + off, o);
            sts_f32(ob + off, prelu_f(o, slope));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar(kAccEmpty + acc));
      fence_proxy_async();
      if (nbuf == 2 && tid == 0) bulk_wait_read();
      named_bar(1, kLtEpiThreads);
      if (tid == 0) {
        if (HAS_ACT && tile + 2 * (int64_t)ncta < ntiles) fetch_zin(tile + 2 * (int64_t)ncta, acc);
        for (int j = 0; j < nob; ++j) {
          tma_store_2d(&tmC, j * kKB, (int)(tile * kTsRows), ob + (uint32_t)j * kTsSlab);
          if (HAS_Z) tma_store_2d(&tmZ, j * kKB, (int)(tile * kTsRows), zb + (uint32_t)j * kTsSlab);
        }
        bulk_commit();
      }
    }
    if (HAS_ACT) {                                         // per-CTA slope-gradient partial, fixed order
      __shared__ float dsl_ts[kLtEpiWarps];
      dsl = warp_sum(dsl);
      if (lane == 0) dsl_ts[warp] = dsl;
      named_bar(1, kLtEpiThreads);
      if (tid == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kLtEpiWarps; ++w) t += dsl_ts[w];
        dslope_part[blockIdx.x] = t;
      }
      if (colsum_part && c < N) colsum_part[((int64_t)blockIdx.x * 2 + rh) * N + c] = csum;
    }
    if (tid == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kLtMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-fed dW: the copy engine brings [32 rows x M] / [32 rows x N] boxes of dY / X (row-major, unswizzled)
// into a raw ring; converter warps transpose + split them smem -> smem into the K-major stages.  A "unit" is
// 32 columns x 4 rows: a lane owns one column, reads it for 4 consecutive rows (4 conflict-free LDS.32) and
// that is exactly one 16-byte K chunk of that column's row in the SWIZZLE_128B K-major layout (consecutive
// lanes -> different swizzled chunks, conflict-free STS.128).  No global-load latency sits on any warp.
// The converters were instruction-bound with generic per-unit index math, so every warp precomputes its <= NU
// units (source / destination offsets) once and the K-block loop is 4 LDS + split + 2 STS per unit.
// 18 warps: 0-3 convert, then run the epilogue; 4 MMA issue + TMEM; 5 TMA producer; 6-17 convert.
constexpr int kDwWarps = 18, kDwThreads = kDwWarps * 32, kDwConv = 16;
#ifndef GCL_DW_GROUPS
#define GCL_DW_GROUPS 2
#endif
constexpr int kDwGroups = GCL_DW_GROUPS;

template <int NU>
__global__ void __launch_bounds__(kDwThreads, 1)
    umma_dw_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       float* __restrict__ part, float* __restrict__ part_colsum, int64_t R, int M, int N, int n_pad,
                       int nst, int nraw, int raw_bytes, int tmem_cols, int64_t rows_per_cta, int wide_b) {
  // wide_b: B_hi and B_lo are adjacent in a stage, so A_hi x [B_hi | B_lo] is ONE MMA of width 2 n_pad (the A_hi
  // slab is read from shared memory once instead of twice); its right half and A_lo x B_hi are the correction
  // terms and get accumulator columns of their own, the epilogue adds the three.
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // the A parts hold only ceil(M / 8) * 8 rows: the M = 128 MMA also reads the following 128 - that rows of
  // whatever comes next in smem, which only produces accumulator rows >= M that nobody reads
  const int a_part = ((M + 7) / 8) * 8 * 128, b_part = n_pad * 128;
  const int stage_bytes = 2 * a_part + 2 * b_part;
  uint8_t* st_base = smem;
  uint8_t* raw = st_base + (size_t)nst * stage_bytes;
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(raw + (size_t)nraw * raw_bytes);
  const uint32_t bars = smem_u32(bar_ptr);
  const int kFull = 0, kEmpty = nst, kRawFull = 2 * nst, kRawEmpty = 2 * nst + nraw, kAccFull = 2 * nst + 2 * nraw;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + kAccFull + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t r_beg = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = min(R, r_beg + rows_per_cta);
  const int64_t nkb = r_end > r_beg ? (r_end - r_beg + kKB - 1) / kKB : 0;

  if (tid == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(bar(kFull + s), kDwConv / kDwGroups);      // one arrival per converter warp of the K-block's group
      mbar_init(bar(kEmpty + s), 1);
    }
    for (int s = 0; s < nraw; ++s) {
      mbar_init(bar(kRawFull + s), 1);
      mbar_init(bar(kRawEmpty + s), kDwConv / kDwGroups);
    }
    mbar_init(bar(kAccFull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps) tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols);
  for (int i = tid; i < nst * stage_bytes / 16; i += kDwThreads)
    reinterpret_cast<float4*>(st_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpiWarps + 1) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int64_t kb = 0; kb < nkb; ++kb) {
        const int row = (int)(r_beg + kb * kKB);
        mbar_wait(bar(kRawEmpty + s), ph);
        mbar_expect_tx(bar(kRawFull + s), (uint32_t)(kKB * (M + N) * 4));
        const uint32_t dst = smem_u32(raw + (size_t)s * raw_bytes);
        tma_load_2d(dst, &tmA, 0, row, bar(kRawFull + s));
        tma_load_2d(dst + (uint32_t)(kKB * M * 4), &tmB, 0, row, bar(kRawFull + s));
        if (++s == nraw) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp != kEpiWarps) {
    // converter warp cw belongs to group cw % kDwGroups, which handles the K-blocks kb % kDwGroups == group:
    // a group may take kDwGroups K-block times for its wait -> load -> split -> store -> fence -> arrive chain
    const int cw = warp < kEpiWarps ? warp : warp - 2;     // 0..15
    const int grp = cw % kDwGroups, wi = cw / kDwGroups;   // wi: 0 .. kDwConv / kDwGroups - 1
    const int na_units = 8 * ((M + 31) / 32), n_units = na_units + 8 * ((N + 31) / 32);
    // this warp's units: byte offsets into a raw stage / a K-major stage (hi part), row pitch, hi -> lo distance
    uint32_t src_off[NU], dst_off[NU], pitch[NU], lo_off[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const int u = wi + i * (kDwConv / kDwGroups);
      const bool isa = u < na_units;
      const int uu = isa ? u : u - na_units;
      const int width = isa ? M : N;
      const int col = (uu >> 3) * 32 + lane, chunk = uu & 7;
      const bool ok = u < n_units && col < width;
      pitch[i] = (uint32_t)width * 4;
      src_off[i] = ok ? (uint32_t)((isa ? 0 : kKB * M * 4) + ((chunk * 4) * width + col) * 4) : 0xffffffffu;
      dst_off[i] = (uint32_t)(isa ? 0 : 2 * a_part) + sw_off(col, chunk);
      lo_off[i] = isa ? (uint32_t)a_part : (uint32_t)b_part;
    }
    // bias gradient = column sums of A: every A unit keeps the sum of the values it converts (its column, its 4 of
    // every 32 rows, its K blocks); the partials land in part_colsum[cta][chunk][group][M] and meet in the reduce
    float cs[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) cs[i] = 0.f;
    const uint32_t st_u32 = smem_u32(st_base), raw_u32 = smem_u32(raw);
    int st = grp % nst, rs = grp % nraw;                  // kDwGroups <= nst, nraw (host guarantees)
    uint32_t ph = 1, rph = 0;
    for (int64_t kb = grp; kb < nkb; kb += kDwGroups) {
      mbar_wait(bar(kRawFull + rs), rph);
      mbar_wait(bar(kEmpty + st), ph);
      const uint32_t rbase = raw_u32 + (uint32_t)rs * (uint32_t)raw_bytes;
      const uint32_t sbase = st_u32 + (uint32_t)st * (uint32_t)stage_bytes;
      float4 v[NU];
#ifndef GCL_DW_SKIP_CONV
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        if (src_off[i] != 0xffffffffu) {
          const uint32_t a = rbase + src_off[i];
          v[i] = make_float4(lds_f32(a), lds_f32(a + pitch[i]), lds_f32(a + 2 * pitch[i]), lds_f32(a + 3 * pitch[i]));
          cs[i] += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
      }
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        if (src_off[i] != 0xffffffffu) {
          float4 h, l;
          split_tf32(v[i].x, h.x, l.x);
          split_tf32(v[i].y, h.y, l.y);
          split_tf32(v[i].z, h.z, l.z);
          split_tf32(v[i].w, h.w, l.w);
          sts_f32x4(sbase + dst_off[i], h);
          sts_f32x4(sbase + dst_off[i] + lo_off[i], l);
        }
      }
#endif
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(kFull + st));
        mbar_arrive(bar(kRawEmpty + rs));
      }
      st += kDwGroups;
      if (st >= nst) { st -= nst; ph ^= 1; }
      rs += kDwGroups;
      if (rs >= nraw) { rs -= nraw; rph ^= 1; }
    }
    if (part_colsum) {
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const int u = wi + i * (kDwConv / kDwGroups);
        const int col = (u >> 3) * 32 + lane, chunk = u & 7;
        if (u < na_units && col < M)
          part_colsum[(((int64_t)blockIdx.x * 8 + chunk) * kDwGroups + grp) * M + col] = cs[i];
      }
    }
    if (warp < kEpiWarps) {
      // ---- epilogue: the accumulator is complete once the last MMA has retired
      const int m = M <= 64 ? (lane < 16 ? warp * 16 + lane : M) : warp * 32 + lane;   // see the MMA warp below
      float* prow = part + ((int64_t)blockIdx.x * M + m) * N;
      if (nkb > 0) {
        mbar_wait(bar(kAccFull), 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < n_pad; c0 += 16) {
          float v16[16];
          tmem_ld16(taddr + c0, v16);
          if (wide_b) {                                  // + (A_hi B_lo) + (A_lo B_hi)
            float c1[16], c2[16];
            tmem_ld16x2(taddr + n_pad + c0, taddr + 2 * n_pad + c0, c1, c2);
#pragma unroll
            for (int j = 0; j < 16; ++j) v16[j] += c1[j] + c2[j];
          }
          if (m < M) {
            if (c0 + 16 <= N && (N & 3) == 0 && al16_dev(part)) {          // 16-byte stores
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(prow + c0 + j) = make_float4(v16[j], v16[j + 1], v16[j + 2], v16[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = c0 + j;
                if (n < N) prow[n] = v16[j];
              }
            }
          }
        }
      } else if (m < M) {
        for (int n = 0; n < N; ++n) prow[n] = 0.f;
      }
    }
  } else {
    // M <= 64: the 64-row MMA shape -- half the A operand reads and half the tensor time; its accumulator row m sits
    // in TMEM lane 32 (m / 16) + m % 16 (16 lanes per sub-partition, cute tmem_frg_1sm "half subpartitions" atom)
    const int mm = M <= 64 ? 64 : kTileM;
    const uint32_t idesc = make_idesc(n_pad, 0, 0, mm), idesc_wide = make_idesc(2 * n_pad, 0, 0, mm);
    const uint32_t st_lo = desc_lo(smem_u32(st_base));
    int st = 0;
    uint32_t ph = 0;
    for (int64_t kb = 0; kb < nkb; ++kb) {
      mbar_wait(bar(kFull + st), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ah = st_lo + (uint32_t)st * ((uint32_t)stage_bytes >> 4), al = ah + ((uint32_t)a_part >> 4);
        const uint32_t bh = al + ((uint32_t)a_part >> 4), bl = bh + ((uint32_t)b_part >> 4);
#ifndef GCL_DW_SKIP_MMA
        if (wide_b) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acc = (kb | ks) ? 1u : 0u;
            umma_tf32(tmem_base + 2 * n_pad, desc_k128(al + 2 * ks), desc_k128(bh + 2 * ks), idesc, acc);
            umma_tf32(tmem_base, desc_k128(ah + 2 * ks), desc_k128(bh + 2 * ks), idesc_wide, acc);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_3x(tmem_base, ah + 2 * ks, al + 2 * ks, bh + 2 * ks, bl + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
        }
#else
        if (kb == 0) umma_3x(tmem_base, ah, al, bh, bl, idesc, 0u);
#endif
        umma_commit(bar(kEmpty + st));
        if (kb == nkb - 1) umma_commit(bar(kAccFull));
      }
      __syncwarp();
      if (++st == nst) { st = 0; ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// dW for wide layers (64 < M <= 128): the dY^T operand goes to TENSOR MEMORY instead of shared memory.
// In umma_dw_tma_kernel both operands are transposed + split smem -> smem and read back by three MMAs: ~208 KB of
// shared-memory traffic per 32-row block (32 KB of HBM data), and ncu shows the 16 converter warps, not HBM or
// the tensor pipe, as the limit (0.62 of the HBM roofline at 128 x 128).  Here a converter thread owns output
// channel m = TMEM lane m: it reads its column of the raw dY block for 16 consecutive rows (conflict-free
// LDS.32), splits and writes hi / lo with two tcgen05.st -- that IS the K-major A operand.  The X side keeps the
// K-major SWIZZLE_128B stages (hi | lo adjacent): A_hi x [B_hi | B_lo] is one MMA of N = 2 n_pad, A_lo x B_hi
// accumulates into the correction columns.  144 KB of shared-memory traffic per block, half the converter work.
// TMEM: columns 0 .. 2 n_pad accumulator (main | correction), 256 + 64 st .. the 4 A stages (32 hi | 32 lo).
// 18 warps: 0-7 dY converters (warp w: lanes 32 (w % 4).., rows 16 (w / 4)..; 0-3 then run the epilogue), 8 MMA
// issue + TMEM allocation, 9 TMA producer, 10-17 X converters.
constexpr int kDtsStages = 4, kDtsA0 = 256, kDtsAWarps = 8, kDtsBWarps = 8;

template <int NUB>
__global__ void __launch_bounds__(kDwThreads, 1)
    umma_dw_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      float* __restrict__ part, float* __restrict__ part_colsum, int64_t R, int M, int N, int n_pad,
                      int nraw, int raw_bytes, int64_t rows_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_part = n_pad * 128, stage_bytes = 2 * b_part;
  uint8_t* st_base = smem;
  uint8_t* raw = st_base + (size_t)kDtsStages * stage_bytes;
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(raw + (size_t)nraw * raw_bytes);
  const uint32_t bars = smem_u32(bar_ptr);
  const int kFull = 0, kEmpty = kDtsStages, kRawFull = 2 * kDtsStages, kRawEmpty = kRawFull + nraw, kAccFull = kRawEmpty + nraw;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + kAccFull + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t r_beg = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = min(R, r_beg + rows_per_cta);
  const int64_t nkb = r_end > r_beg ? (r_end - r_beg + kKB - 1) / kKB : 0;
  constexpr int kMmaWarp = kDtsAWarps, kTmaWarp = kDtsAWarps + 1;

  if (tid == 0) {
    for (int s = 0; s < kDtsStages; ++s) {
      mbar_init(bar(kFull + s), kDtsAWarps + kDtsBWarps);      // one arrival per converter warp
      mbar_init(bar(kEmpty + s), 1);
    }
    for (int s = 0; s < nraw; ++s) {
      mbar_init(bar(kRawFull + s), 1);
      mbar_init(bar(kRawEmpty + s), kDtsAWarps + kDtsBWarps);
    }
    mbar_init(bar(kAccFull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512u);
  for (int i = tid; i < kDtsStages * stage_bytes / 16; i += kDwThreads)
    reinterpret_cast<float4*>(st_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t raw_u32 = smem_u32(raw), st_u32 = smem_u32(st_base);

  if (warp == kTmaWarp) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      for (int64_t kb = 0; kb < nkb; ++kb) {
        const int row = (int)(r_beg + kb * kKB);
        mbar_wait(bar(kRawEmpty + s), ph);
        mbar_expect_tx(bar(kRawFull + s), (uint32_t)(kKB * (M + N) * 4));
        const uint32_t dst = raw_u32 + (uint32_t)s * (uint32_t)raw_bytes;
        tma_load_2d(dst, &tmA, 0, row, bar(kRawFull + s));
        tma_load_2d(dst + (uint32_t)(kKB * M * 4), &tmB, 0, row, bar(kRawFull + s));
        if (++s == nraw) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp < kDtsAWarps) {
    // ---- dY converters: raw [32 rows][M] -> TMEM lane m, columns = rows of the block (hi | lo)
    const int q = warp & 3, hr = warp >> 2;
    const int m = q * 32 + lane;
    const bool m_ok = m < M;
    const uint32_t pitch = (uint32_t)M * 4u;
    const uint32_t src = (uint32_t)((16 * hr) * M + (m_ok ? m : 0)) * 4u;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kDtsA0 + 16 * hr);
    float cs = 0.f;
    int st = 0, rs = 0;
    uint32_t ph = 1, rph = 0;
    for (int64_t kb = 0; kb < nkb; ++kb) {
      mbar_wait(bar(kRawFull + rs), rph);
      mbar_wait(bar(kEmpty + st), ph);
      tc_fence_after();
      const uint32_t a = raw_u32 + (uint32_t)rs * (uint32_t)raw_bytes + src;
      float v[16], hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = lds_f32(a + (uint32_t)j * pitch);
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (!m_ok) v[j] = 0.f;
        t += v[j];
        split_tf32(v[j], hi[j], lo[j]);
      }
      cs += t;
      tmem_st16(tl + (uint32_t)(st * 64), hi);
      tmem_st16(tl + (uint32_t)(st * 64 + 32), lo);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(kFull + st));
        mbar_arrive(bar(kRawEmpty + rs));
      }
      if (++st == kDtsStages) { st = 0; ph ^= 1; }
      if (++rs == nraw) { rs = 0; rph ^= 1; }
    }
    if (part_colsum && m_ok) part_colsum[((int64_t)blockIdx.x * 2 + hr) * M + m] = cs;
    if (warp < 4) {
      // ---- epilogue: the accumulator is complete once the last MMA has retired
      float* prow = part + ((int64_t)blockIdx.x * M + m) * N;
      if (nkb > 0) {
        mbar_wait(bar(kAccFull), 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < n_pad; c0 += 16) {
          float v16[16], c16[16];
          tmem_ld16x2(taddr + c0, taddr + n_pad + c0, v16, c16);
          if (m_ok) {
            if (c0 + 16 <= N && (N & 3) == 0 && al16_dev(part)) {          // 16-byte stores
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(prow + c0 + j) = make_float4(v16[j] + c16[j], v16[j + 1] + c16[j + 1],
                                                                        v16[j + 2] + c16[j + 2], v16[j + 3] + c16[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = c0 + j;
                if (n < N) prow[n] = v16[j] + c16[j];
              }
            }
          }
        }
      } else if (m_ok) {
        for (int n = 0; n < N; ++n) prow[n] = 0.f;
      }
    }
  } else if (warp > kTmaWarp) {
    // ---- X converters: (32 columns x 4 rows) units, transpose + split into the K-major stage
    const int bw = warp - (kTmaWarp + 1);                 // 0..7
    const int n_units = 8 * ((N + 31) / 32);
    uint32_t src_off[NUB], dst_off[NUB];
#pragma unroll
    for (int i = 0; i < NUB; ++i) {
      const int u = bw + i * kDtsBWarps;
      const int col = (u >> 3) * 32 + lane, chunk = u & 7;
      const bool ok = u < n_units && col < N;
      src_off[i] = ok ? (uint32_t)(kKB * M * 4 + ((chunk * 4) * N + col) * 4) : 0xffffffffu;
      dst_off[i] = sw_off(col, chunk);
    }
    const uint32_t pitch = (uint32_t)N * 4u;
    int st = 0, rs = 0;
    uint32_t ph = 1, rph = 0;
    for (int64_t kb = 0; kb < nkb; ++kb) {
      mbar_wait(bar(kRawFull + rs), rph);
      mbar_wait(bar(kEmpty + st), ph);
      const uint32_t rbase = raw_u32 + (uint32_t)rs * (uint32_t)raw_bytes;
      const uint32_t sbase = st_u32 + (uint32_t)st * (uint32_t)stage_bytes;
      float4 v[NUB];
#pragma unroll
      for (int i = 0; i < NUB; ++i) {
        if (src_off[i] != 0xffffffffu) {
          const uint32_t a = rbase + src_off[i];
          v[i] = make_float4(lds_f32(a), lds_f32(a + pitch), lds_f32(a + 2 * pitch), lds_f32(a + 3 * pitch));
        }
      }
#pragma unroll
      for (int i = 0; i < NUB; ++i) {
        if (src_off[i] != 0xffffffffu) {
          float4 h, l;
          split_tf32(v[i].x, h.x, l.x);
          split_tf32(v[i].y, h.y, l.y);
          split_tf32(v[i].z, h.z, l.z);
          split_tf32(v[i].w, h.w, l.w);
          sts_f32x4(sbase + dst_off[i], h);
          sts_f32x4(sbase + dst_off[i] + (uint32_t)b_part, l);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(kFull + st));
        mbar_arrive(bar(kRawEmpty + rs));
      }
      if (++st == kDtsStages) { st = 0; ph ^= 1; }
      if (++rs == nraw) { rs = 0; rph ^= 1; }
    }
  } else {
    // ---- MMA issuer
    const uint32_t idesc = make_idesc(n_pad, 0, 0), idesc_wide = make_idesc(2 * n_pad, 0, 0);
    const uint32_t st_lo = desc_lo(st_u32);
    int st = 0;
    uint32_t ph = 0;
    for (int64_t kb = 0; kb < nkb; ++kb) {
      mbar_wait(bar(kFull + st), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b = st_lo + (uint32_t)st * ((uint32_t)stage_bytes >> 4);
        const uint32_t a_hi = tmem_base + (uint32_t)(kDtsA0 + st * 64), a_lo = a_hi + 32u;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_tf32_ts(tmem_base, a_hi + 8 * ks, desc_k128(b + 2 * ks), idesc_wide, (kb | ks) ? 1u : 0u);
          umma_tf32_ts(tmem_base + (uint32_t)n_pad, a_lo + 8 * ks, desc_k128(b + 2 * ks), idesc, 1u);
        }
        umma_commit(bar(kEmpty + st));
        if (kb == nkb - 1) umma_commit(bar(kAccFull));
      }
      __syncwarp();
      if (++st == kDtsStages) { st = 0; ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
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


namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// the driver's tensor-map encoder, resolved through the runtime (no link-time dependency on libcuda)
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
// fp32 [rows, cols] row-major, boxes of [box_rows x box_cols]; out-of-range elements load as 0 / are not stored
bool make_map_2d(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int box_rows, int box_cols,
                 CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct TmaLinPlan {
  int n_pad, nkb, nob, s_raw, s_lo, nbuf, tmem_cols;
  size_t smem;
  bool ok;
};
TmaLinPlan plan_linear_tma(int64_t N, int64_t K, bool has_z, bool has_act = false) {
  TmaLinPlan p{};
  p.n_pad = (int)((N + 15) / 16 * 16);
  p.nkb = (int)((K + kKB - 1) / kKB);
  p.nob = (int)((N + kKB - 1) / kKB);
  // narrow layers (the 19/20-channel output head): boxes are wider than the tensor, the tensor map zero-fills loads
  // and clips stores
  static const int min_w = getenv("GCL_TMA_MIN_WIDTH") ? atoi(getenv("GCL_TMA_MIN_WIDTH")) : 8;
  if (N < min_w || K < min_w || p.n_pad > 256 || (N & 3) || (K & 3)) return p;
  const long w_bytes = 2L * p.nkb * p.n_pad * 128;
  const long fixed = 1024 + 512 + 4L * p.n_pad * 3 + 4L * 2 * 2 * kTileM * 2;   // bias, att, score partials
  long slabs = 0, out_bytes = 0;
  for (p.nbuf = 2; p.nbuf >= 1; --p.nbuf) {       // two output buffers if at least 7 operand slabs remain
    out_bytes = (long)p.nbuf * p.nob * kSlabBytes * (has_z ? 2 : 1) + (has_act ? (long)p.nob * kSlabBytes : 0);
    slabs = ((long)kMaxSmem - w_bytes - out_bytes - fixed) / kSlabBytes;
    if (slabs >= (p.nbuf == 2 ? 7 : 5)) break;
  }
  if (p.nbuf < 1) return p;
  p.s_lo = (int)(slabs * 2 / 5);
  if (p.s_lo > 4) p.s_lo = 4;
  p.s_raw = (int)(slabs - p.s_lo);
  if (p.s_raw > 8) p.s_raw = 8;
  p.smem = (size_t)(w_bytes + out_bytes + fixed + (long)(p.s_raw + p.s_lo) * kSlabBytes);
  int cols = 32;
  while (cols < 4 * p.n_pad) cols <<= 1;     // two tiles in flight x (main + correction accumulator)
  if (cols > 512) return p;
  p.tmem_cols = cols;
  p.ok = true;
  return p;
}

// Weights-stationary kernel for 64 < N <= 128, K <= 128 (GCL_UMMA_NO_TS=1 switches it off for A/B runs).
static const bool g_no_ts = [] {
  const char* e = getenv("GCL_UMMA_NO_TS");
  return e && e[0] == '1';
}();
int umma_linear_ts(const float* A, const float* W_nk, float* C, int64_t M, int64_t N, int64_t K, const float* bias,
                   const float* slope, float* z_out, const float* z_in, const float* act_slope, float* dslope_part,
                   int* n_parts, cudaStream_t s, float* colsum_part, int* n_colsum_parts) {
  // <= 64 output channels use half of the 128 TMEM lanes, i.e. 4 of the 8 epilogue warps (a warp reads only its own
  // lane quarter): measured worth it only for the variant that also stores the pre-activation (0.92 vs 0.73 at
  // 64 -> 64); the PReLU'-epilogue variant is epilogue-bound there (0.47 vs 0.68) and stays on the row-major kernel
  static const bool no_z64 = getenv("GCL_TS_NO_Z64") && getenv("GCL_TS_NO_Z64")[0] == '1';     // A/B switch
  // (... and the PReLU'-epilogue variant when the caller also wants the column sums, which only this kernel gets for
  // free: 256 vs 318 us at R = 1 376 272, 64 -> 64; without them the row-major kernel is level, 237 vs 248 us)
  const int min_n = ((z_out && !act_slope && !no_z64) || (act_slope && colsum_part && !no_z64)) ? kKB : 65;
  if (g_no_ts || N < min_n || N > 128 || K < kKB || K > kTsWCols || (N & 3) || (K & 3)) return GCL_ERR_UNSUPPORTED;
  const int has_z = z_out ? 1 : 0, has_act = act_slope ? 1 : 0;
  const int nkb = (int)((K + kKB - 1) / kKB), nob = (int)((N + kKB - 1) / kKB);
  const long fixed = 1024 + 8L * 64;
  int nbuf = 2, nst = 0;
  for (; nbuf >= 1; --nbuf) {      // two output buffers if at least 6 operand stages remain
    const long out_bytes = (long)nbuf * nob * kTsSlab * (1 + has_z) + 2L * has_act * nob * kTsSlab;
    nst = (int)(((long)kMaxSmem - out_bytes - fixed) / kTsStage);
    if (nst >= (nbuf == 2 ? 6 : 4)) break;
  }
  if (nbuf < 1) return GCL_ERR_UNSUPPORTED;
  if (nst > 10) nst = 10;
  const size_t smem = (size_t)fixed + (size_t)nst * kTsStage + (size_t)nbuf * nob * kTsSlab * (1 + has_z) +
                      2 * (size_t)has_act * nob * kTsSlab;
  CUtensorMap tmA, tmC, tmZ, tmZin;
  if (!make_map_2d(&tmA, A, M, K, kTsRows, kKB, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !make_map_2d(&tmC, C, M, N, kTsRows, kKB, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !make_map_2d(&tmZ, z_out ? z_out : C, M, N, kTsRows, kKB, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !make_map_2d(&tmZin, z_in ? z_in : C, M, N, kTsRows, kKB, CU_TENSOR_MAP_SWIZZLE_128B))
    return GCL_ERR_UNSUPPORTED;
  const int64_t ntiles = (M + kTsRows - 1) / kTsRows;
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kTsThreads, smem, s>>>(tmA, tmC, tmZ, tmZin, W_nk, M, (int)N, (int)K, nkb, nob, nst, nbuf, bias, slope,
                                        act_slope, dslope_part, has_act ? colsum_part : nullptr);
    return cudaSuccess;
  };
  const cudaError_t e = has_act ? launch(umma_linear_ts_kernel<false, true>)
                        : has_z ? launch(umma_linear_ts_kernel<true, false>)
                                : launch(umma_linear_ts_kernel<false, false>);
  if (e != cudaSuccess) return fail_cuda(e, "umma_linear_ts(smem attr)");
  if (n_parts) *n_parts = grid;
  if (n_colsum_parts) *n_colsum_parts = (has_act && colsum_part) ? 2 * grid : 0;
  GCL_CHECK_LAUNCH("umma_linear_ts");
  return GCL_OK;
}

int umma_linear_tma(const float* A, const float* W_nk, float* C, int64_t M, int64_t N, int64_t K, const float* bias,
                    const float* slope, float* z_out, const float* z_in, const float* act_slope, float* dslope_part,
                    int* n_parts, const float* att, float* sc_src, float* sc_dst, cudaStream_t s,
                    float* colsum_part = nullptr, int* n_colsum_parts = nullptr) {
  if (n_colsum_parts) *n_colsum_parts = 0;
  if (act_slope && (!z_in || !dslope_part || !al16(z_in) || z_out)) return GCL_ERR_UNSUPPORTED;
  if (!(al16(A) && al16(W_nk) && al16(C) && (!z_out || al16(z_out))) || M <= 0 || M > 0x7fffff00LL)
    return GCL_ERR_UNSUPPORTED;
  if (!att) {
    const int rc = umma_linear_ts(A, W_nk, C, M, N, K, bias, slope, z_out, z_in, act_slope, dslope_part, n_parts, s,
                                  colsum_part, n_colsum_parts);
    if (rc != GCL_ERR_UNSUPPORTED) return rc;
  }
  TmaLinPlan p = plan_linear_tma(N, K, z_out != nullptr, act_slope != nullptr);
  int n_split = 1;
  if (!p.ok && N % 64 == 0) {            // too wide for one CTA's smem: two CTAs per row tile, half the columns each
    p = plan_linear_tma(N / 2, K, z_out != nullptr, act_slope != nullptr);
    n_split = 2;
  }
  if (!p.ok || (att && n_split != 1)) return GCL_ERR_UNSUPPORTED;
  CUtensorMap tmA, tmC, tmZ, tmZin;
  if (!make_map_2d(&tmA, A, M, K, kTileM, kKB, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !make_map_2d(&tmC, C, M, N, kTileM, kKB, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !make_map_2d(&tmZ, z_out ? z_out : C, M, N, kTileM, kKB, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !make_map_2d(&tmZin, z_in ? z_in : C, M, N, kTileM, kKB, CU_TENSOR_MAP_SWIZZLE_128B))
    return GCL_ERR_UNSUPPORTED;
  const int64_t ntiles = (M + kTileM - 1) / kTileM;
  const int64_t per = kNumSMs / n_split;
  const int grid = (int)(ntiles < per ? ntiles : per) * n_split;
  cudaError_t e =
      cudaFuncSetAttribute(umma_linear_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return fail_cuda(e, "umma_linear_tma(smem attr)");
  umma_linear_tma_kernel<<<grid, kLtThreads, p.smem, s>>>(tmA, tmC, tmZ, W_nk, M, (int)(N / n_split), (int)K, p.n_pad,
                                                          p.nkb, p.nob, p.s_raw, p.s_lo, p.nbuf, p.tmem_cols,
                                                          z_out ? 1 : 0, bias, slope, n_split, tmZin, act_slope,
                                                          dslope_part, att, sc_src, sc_dst);
  if (n_parts) *n_parts = grid;
  GCL_CHECK_LAUNCH("umma_linear_tma");
  return GCL_OK;
}
}  // namespace

// Returns GCL_OK when launched, GCL_ERR_UNSUPPORTED when the shape does not fit (caller falls back to
// the FFMA kernel), or an error.
int umma_linear(const float* A, const float* W_nk, float* C, int64_t M, int64_t N, int64_t K, const float* bias,
                const float* slope, float* z_out, cudaStream_t s, const float* z_in, const float* act_slope,
                float* dslope_part, int* n_parts, const float* att, float* sc_src, float* sc_dst, float* colsum_part,
                int* n_colsum_parts) {
  if (n_colsum_parts) *n_colsum_parts = 0;
  if (act_slope || att) {   // only the TMA kernel has the fused PReLU-backward / attention-score epilogues
    if (g_force_register_staging) return GCL_ERR_UNSUPPORTED;
    return umma_linear_tma(A, W_nk, C, M, N, K, bias, slope, z_out, z_in, act_slope, dslope_part, n_parts, att, sc_src,
                           sc_dst, s, colsum_part, n_colsum_parts);
  }
  if (!g_force_register_staging) {
    const int rc = umma_linear_tma(A, W_nk, C, M, N, K, bias, slope, z_out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   nullptr, s);
    if (rc != GCL_ERR_UNSUPPORTED) return rc;
  }
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
            cudaStream_t s, int* n_bias_parts) {
  DwUmmaPlan p = plan_dw(R, M, N);
  if (!p.ok) return GCL_ERR_UNSUPPORTED;
  if (n_bias_parts) *n_bias_parts = p.grid;
  if (!g_force_register_staging && (M & 3) == 0 && (N & 3) == 0 && al16(A) && al16(B) && R <= 0x7fffff00LL) {
    const int raw_bytes = (int)((kKB * (M + N) * 4 + 127) / 128 * 128);
    static const bool no_dw_ts = getenv("GCL_UMMA_NO_TS") && getenv("GCL_UMMA_NO_TS")[0] == '1';
    if (!no_dw_ts && M > 64) {          // wide layers: dY^T in tensor memory
      const int n_pad = (int)((N + 15) / 16 * 16);
      const size_t stage = 2 * (size_t)n_pad * 128, fixed = 1024 + 512;
      long nraw = ((long)kMaxSmem - (long)fixed - (long)kDtsStages * (long)stage) / raw_bytes;
      if (nraw > 10) nraw = 10;
      CUtensorMap tmA, tmB;
      if (nraw >= 2 && make_map_2d(&tmA, A, R, M, kKB, (int)M, CU_TENSOR_MAP_SWIZZLE_NONE) &&
          make_map_2d(&tmB, B, R, N, kKB, (int)N, CU_TENSOR_MAP_SWIZZLE_NONE)) {
        const size_t smem = fixed + (size_t)kDtsStages * stage + (size_t)nraw * raw_bytes;
        const int nub = (8 * (int)((N + 31) / 32) + kDtsBWarps - 1) / kDtsBWarps;       // 1..4
        auto launch = [&](auto kern) -> int {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return fail_cuda(e, "umma_dw_ts(smem attr)");
          kern<<<p.grid, kDwThreads, smem, s>>>(tmA, tmB, part, part_colsum, R, (int)M, (int)N, n_pad, (int)nraw, raw_bytes,
                                                p.rows_per_cta);
          return GCL_OK;
        };
        const int rc = nub <= 2 ? launch(umma_dw_ts_kernel<2>) : launch(umma_dw_ts_kernel<4>);
        if (rc != GCL_OK) return rc;
        GCL_CHECK_LAUNCH("umma_dw_ts");
        if (n_bias_parts) *n_bias_parts = p.grid * 2;
        return GCL_OK;
      }
    }
    // raw ring + as many K-major stages as fit
    const int n_pad = (int)((N + 15) / 16 * 16);           // no ones-row here: the converters sum the columns of A
    const size_t stage = 2 * (size_t)((M + 7) / 8 * 8) * 128 + 2 * (size_t)n_pad * 128;
    const size_t fixed = 1024 + 512;
    static const int kDwMinRaw = getenv("GCL_DW_MINRAW") ? atoi(getenv("GCL_DW_MINRAW")) : 6;
    int nst = 4, nraw = 0;
    for (; nst >= 2; --nst) {
      const long room = (long)kMaxSmem - (long)fixed - (long)nst * (long)stage;
      nraw = room > 0 ? (int)(room / raw_bytes) : 0;
      if (nraw >= kDwMinRaw || (nst == 2 && nraw >= 2)) break;
    }
    if (nraw > 10) nraw = 10;
    CUtensorMap tmA, tmB;
    if (nst >= kDwGroups && nraw >= kDwGroups && nst >= 2 && nraw >= 2 && (size_t)nraw * raw_bytes >= (size_t)kPartBytes && make_map_2d(&tmA, A, R, M, kKB, (int)M, CU_TENSOR_MAP_SWIZZLE_NONE) &&
        make_map_2d(&tmB, B, R, N, kKB, (int)N, CU_TENSOR_MAP_SWIZZLE_NONE)) {
      const size_t smem = fixed + (size_t)nst * stage + (size_t)nraw * raw_bytes;
      static const bool no_wide = getenv("GCL_DW_NO_WIDE") && getenv("GCL_DW_NO_WIDE")[0] == '1';
      const int wide_b = (!no_wide && 2 * n_pad <= 256) ? 1 : 0;     // UMMA N <= 256
      int wide_cols = 32;
      while (wide_cols < (wide_b ? 3 : 1) * n_pad) wide_cols <<= 1;
      const int per_grp = kDwConv / kDwGroups;
      const int n_units = 8 * (int)((M + 31) / 32 + (N + 31) / 32), nu = (n_units + per_grp - 1) / per_grp;
      auto launch = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail_cuda(e, "umma_dw_tma(smem attr)");
        kern<<<p.grid, kDwThreads, smem, s>>>(tmA, tmB, part, part_colsum, R, (int)M, (int)N, n_pad, nst, nraw,
                                              raw_bytes, wide_cols, p.rows_per_cta, wide_b);
        return GCL_OK;
      };
      int rc;
      if (nu <= 2) rc = launch(umma_dw_tma_kernel<2>);
      else if (nu <= 4) rc = launch(umma_dw_tma_kernel<4>);
      else if (nu <= 6) rc = launch(umma_dw_tma_kernel<6>);
      else if (nu <= 8) rc = launch(umma_dw_tma_kernel<8>);
      else rc = launch(umma_dw_tma_kernel<16>);
      if (rc != GCL_OK) return rc;
      GCL_CHECK_LAUNCH("umma_dw_tma");
      if (n_bias_parts) *n_bias_parts = p.grid * 8 * kDwGroups;
      return GCL_OK;
    }
  }
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

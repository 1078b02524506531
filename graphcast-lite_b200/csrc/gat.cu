// K4/K5: GATConv / SparseGATConv message passing, forward and backward, fp32, no atomics.
//
// Replaces PyG GATConv.edge_update + softmax + propagate (+ head mean / bias), which the reference
// uses at /root/reference/src/models.py:336-357 (GATConv) and :112-151 (SparseGATConv):
//   e_ij   = LeakyReLU(a_src[j] + a_dst[i])                      (j -> i, self loops included)
//   alpha  = exp(e - max_i) / (sum_i exp(.) + 1e-16)             (torch_geometric.utils.softmax)
//   out_i  = mean_h / concat_h ( sum_j alpha_ijh z_jh ) + bias
// Mapping (all three aggregate kernels): a group of L lanes (L = 4..32 128-bit words of a head slice) owns one
// (receiver or sender row, SB samples), 32/L groups share a warp; lane k of the group also holds edge k of the row,
// so per-row scalars (logits, softmax, alpha) live one edge per lane and are broadcast by group shuffles while
// every lane gathers its word of each neighbour row for SB samples.
//   heads == 1 forward : gat_alpha_kernel (coefficients, thread per (sample, row)) + the SpMM kernel with
//                        per-sample weights (spmm.cu), bias and optionally the layer's PReLU in its epilogue
//   heads  > 1 forward : gat_fwd_kernel (fused logits / softmax / aggregate / head mean or concat)
//   backward           : gat_bwd_dst_kernel (d alpha by a butterfly transpose-reduce of the edge dot products,
//                        softmax backward, da_dst) then gat_bwd_src_kernel on the sender-grouped CSR (dz, da_src)
// alpha is written once in CSR order (kept for backward) and optionally scattered to PyG edge order
// (return_attention_weights / pruning).
#include "common.cuh"
#include "scan.cuh"

#ifndef GCL_GAT_UNROLL
#define GCL_GAT_UNROLL 2
#endif
namespace gcl {
namespace {
constexpr int kGatUnroll = GCL_GAT_UNROLL;   // neighbours whose gathers are issued back to back


constexpr int kWarps = 8;

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : slope * v; }

template <int VW>
struct W;
template <>
struct W<4> {
  using T = float4;
  static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ T load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void store(float* p, T v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ void fma(T& a, float w, T v) {
    a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
  }
  static __device__ __forceinline__ T add(T a, T b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
  static __device__ __forceinline__ T scale(T a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
  static __device__ __forceinline__ float dot(T a, T b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
  static __device__ __forceinline__ T xor_add(T a, int o) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
    a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
    return a;
  }
};
template <>
struct W<1> {
  using T = float;
  static __device__ __forceinline__ T zero() { return 0.f; }
  static __device__ __forceinline__ T load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void store(float* p, T v) { *p = v; }
  static __device__ __forceinline__ void fma(T& a, float w, T v) { a = fmaf(w, v, a); }
  static __device__ __forceinline__ T add(T a, T b) { return a + b; }
  static __device__ __forceinline__ T scale(T a, float s) { return a * s; }
  static __device__ __forceinline__ float dot(T a, T b) { return a * b; }
  static __device__ __forceinline__ T xor_add(T a, int o) { return a + __shfl_xor_sync(0xffffffffu, a, o); }
};

// a_src[row,h] = <z[row,h,:], att_src[h,:]>, same for dst.  8 lanes per (row, head), 128-bit loads.
__global__ void gat_scores_kernel(const float* __restrict__ z, const float* __restrict__ att_src,
                                  const float* __restrict__ att_dst, float* __restrict__ a_src,
                                  float* __restrict__ a_dst, int64_t rows, int H, int C) {
  const int64_t item = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;   // (row, head)
  const int gl = threadIdx.x & 7;
  const unsigned mask = 0xffu << ((threadIdx.x & 31) & ~7);
  if (item >= rows * H) return;
  const int h = (int)(item % H);
  const float* zr = z + item * C;
  const float* as = att_src + (int64_t)h * C;
  const float* ad = att_dst + (int64_t)h * C;
  float s = 0.f, d = 0.f;
  if ((C & 3) == 0 && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(att_src) |
                        reinterpret_cast<uintptr_t>(att_dst)) & 15) == 0) {
    for (int c = gl * 4; c < C; c += 32) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(zr + c));
      const float4 p = __ldg(reinterpret_cast<const float4*>(as + c));
      const float4 q = __ldg(reinterpret_cast<const float4*>(ad + c));
      s += v.x * p.x + v.y * p.y + v.z * p.z + v.w * p.w;
      d += v.x * q.x + v.y * q.y + v.z * q.z + v.w * q.w;
    }
  } else {
    for (int c = gl; c < C; c += 8) {
      const float v = zr[c];
      s = fmaf(v, __ldg(as + c), s);
      d = fmaf(v, __ldg(ad + c), d);
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    s += __shfl_xor_sync(mask, s, o, 8);
    d += __shfl_xor_sync(mask, d, o, 8);
  }
  if (gl == 0) {
    a_src[item] = s;
    a_dst[item] = d;
  }
}

// ---- shared structure of the three attention kernels -------------------------------------------------
// A group of L lanes (L = 4..32 >= words per head slice) owns one (row, SB consecutive samples); 32/L
// groups share a warp and use per-group shuffle masks.  The row's neighbour ids are fetched once and
// reused for all SB samples and heads; every edge issues SB independent 128-bit gathers per lane, which
// is what keeps enough bytes in flight (the per-sample-per-warp version of these kernels ran at ~10% of
// the HBM roofline, latency-bound).
template <int L>
__device__ __forceinline__ unsigned group_mask(int lane) {
  return (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << ((lane / L) * L));
}
template <int L>
__device__ __forceinline__ float gsum(float v, unsigned mask) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, L);
  return v;
}
template <int L>
__device__ __forceinline__ float gmax(float v, unsigned mask) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o, L));
  return v;
}

// Hot-loop rules for the three aggregate kernels below (from the SASS of the first version, which spent ~85% of
// its instructions on 64-bit address arithmetic, per-sample `s < nb` branches around every gather and spills):
//   * a ragged last sample block re-reads sample nb - 1 instead of branching (only stores are guarded), idle
//     lanes (C / VW < L) re-read word 0;
//   * per-sample base pointers are formed once per (row, head); a gather is then base + (uint32) c * HC;
//   * 2 CTAs / SM (128 registers) -- the gathers of SB samples x 2 unrolled neighbours give each lane 8 independent
//     128-bit loads in flight, which covers the L2 latency without a third CTA.
template <int VW, int L, int SB>
__global__ void __launch_bounds__(kWarps * 32, SB >= 4 ? 2 : 4)
    gat_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                   const int32_t* __restrict__ perm, const float* __restrict__ z, const float* __restrict__ a_src,
                   const float* __restrict__ a_dst, const float* __restrict__ bias, float* __restrict__ out,
                   float* __restrict__ alpha_csr, float* __restrict__ alpha_pyg, int64_t N, int64_t nnz, int B, int H,
                   int C, int concat, float slope) {
  using V = W<VW>;
  constexpr int kGroups = 32 / L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane & (L - 1);
  const unsigned mask = group_mask<L>(lane);
  const int64_t i = ((int64_t)blockIdx.x * kWarps + warp) * kGroups + lane / L;
  if (i >= N) return;
  const int b0 = blockIdx.y * SB;
  const int nb = min(SB, B - b0);
  const int32_t beg = rowptr[i], end = rowptr[i + 1];
  const int HC = H * C, Cout = concat ? HC : C;
  const int words = (C + VW - 1) / VW, nchunks = (words + L - 1) / L;
  const float hscale = concat ? 1.f : 1.f / (float)H;     // head mean folded into the attention weight
  const bool single = (end - beg) <= L;                   // whole row in one pass of the group's lanes
  int64_t srow[SB];                                       // first node row of (clamped) sample s
#pragma unroll
  for (int s = 0; s < SB; ++s) srow[s] = (int64_t)(b0 + min(s, nb - 1)) * N;

  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int off = (chunk * L + gl) * VW;
    const bool live = off < C;
    const int offc = live ? off : 0;
    typename V::T acc[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) acc[s] = V::zero();
    for (int h = 0; h < H; ++h) {
      const float* zp[SB];                                 // z[b, 0, h, offc]
      const float* asp[SB];                                // a_src[b, 0, h]
      float adi[SB], m[SB], rl[SB];
#pragma unroll
      for (int s = 0; s < SB; ++s) {
        zp[s] = z + srow[s] * HC + h * C + offc;
        asp[s] = a_src + srow[s] * H + h;
        adi[s] = a_dst[(srow[s] + i) * H + h];
      }
      float e_own[SB];                                     // single-pass rows: this lane's edge logit
      const int32_t k_own = beg + gl;
      const bool own = single && k_own < end;
      const int32_t c_own = own ? col[k_own] : 0;
      if (single) {
        // lane k holds edge k: max and sum are two group reductions, one exp per edge
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          e_own[s] = own ? leaky(asp[s][(uint32_t)c_own * (uint32_t)H] + adi[s], slope) : -INFINITY;
          m[s] = gmax<L>(e_own[s], mask);
          e_own[s] = own ? __expf(e_own[s] - m[s]) : 0.f;
          rl[s] = 1.f / (gsum<L>(e_own[s], mask) + 1e-16f);
        }
      } else {
        float l[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s) { m[s] = -INFINITY; l[s] = 0.f; }
        for (int32_t k = beg + gl; k < end; k += L) {        // online (max, sum) per sample over long rows
          const uint32_t ch = (uint32_t)col[k] * (uint32_t)H;
#pragma unroll
          for (int s = 0; s < SB; ++s) {
            const float e = leaky(asp[s][ch] + adi[s], slope);
            const float mn = fmaxf(m[s], e);
            l[s] = l[s] * __expf(m[s] - mn) + __expf(e - mn);
            m[s] = mn;
          }
        }
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          const float mg = gmax<L>(m[s], mask);
          rl[s] = 1.f / (gsum<L>(m[s] == -INFINITY ? 0.f : l[s] * __expf(m[s] - mg), mask) + 1e-16f);
          m[s] = mg;
        }
      }
      for (int32_t base = beg; base < end; base += L) {
        const int32_t k = base + gl;
        const int n = min(L, end - base);
        int32_t c_reg = 0;
        float a_reg[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s) a_reg[s] = 0.f;
        if (k < end) {
          c_reg = single ? c_own : col[k];
          const int32_t pk = (alpha_pyg && chunk == 0) ? perm[k] : 0;
#pragma unroll
          for (int s = 0; s < SB; ++s) {
            float p;
            if (single) p = e_own[s];
            else p = __expf(leaky(asp[s][(uint32_t)c_reg * (uint32_t)H] + adi[s], slope) - m[s]);
            const float al = p * rl[s];
            a_reg[s] = al * hscale;
            if (chunk == 0 && s < nb) {
              alpha_csr[((int64_t)(b0 + s) * nnz + k) * H + h] = al;
              if (alpha_pyg) alpha_pyg[((int64_t)(b0 + s) * nnz + pk) * H + h] = al;
            }
          }
        }
#pragma unroll kGatUnroll
        for (int j = 0; j < n; ++j) {
          const uint32_t ro = (uint32_t)__shfl_sync(mask, c_reg, j, L) * (uint32_t)HC;
          float a[SB];
          typename V::T v[SB];
#pragma unroll
          for (int s = 0; s < SB; ++s) {
            a[s] = __shfl_sync(mask, a_reg[s], j, L);
            v[s] = V::load(zp[s] + ro);
          }
#pragma unroll
          for (int s = 0; s < SB; ++s) V::fma(acc[s], a[s], v[s]);
        }
      }
      if (concat) {
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          if (live && s < nb) {
            typename V::T o = acc[s];
            if (bias) o = V::add(o, V::load(bias + h * C + off));
            V::store(out + ((int64_t)(b0 + s) * N + i) * Cout + h * C + off, o);
          }
          acc[s] = V::zero();
        }
      }
    }
    if (!concat && live) {
#pragma unroll
      for (int s = 0; s < SB; ++s) {
        if (s < nb) {
          typename V::T o = acc[s];
          if (bias) o = V::add(o, V::load(bias + off));
          V::store(out + ((int64_t)(b0 + s) * N + i) * Cout + off, o);
        }
      }
    }
  }
}

// Attention coefficients alone, one thread per (sample, receiver, head): alpha = softmax_k LeakyReLU(a_src[col_k] +
// a_dst[i]).  With the coefficients in memory (the backward needs them there anyway) the single-head aggregation
// IS the SpMM kernel with per-sample weights: no exp, no reductions and a third of the instructions on the
// gather path (the fused kernel above spends ~70% of its instructions outside the gathers).
__global__ void __launch_bounds__(256)
    gat_alpha_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const int32_t* __restrict__ perm, const float* __restrict__ a_src, const float* __restrict__ a_dst,
                     float* __restrict__ alpha_csr, float* __restrict__ alpha_pyg, int64_t N, int64_t nnz, int B, int H,
                     float slope) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;     // ((b * N) + i) * H + h, i.e. the a_dst index
  if (t >= (int64_t)B * N * H) return;
  const int h = (int)(t % H);
  const int64_t bi = t / H, i = bi % N, b = bi / N;
  const int32_t beg = rowptr[i], end = rowptr[i + 1];
  const float adi = a_dst[t];
  const float* as = a_src + b * N * H + h;
  float m = -INFINITY;
  for (int32_t k = beg; k < end; ++k) m = fmaxf(m, leaky(as[(uint32_t)col[k] * (uint32_t)H] + adi, slope));
  float l = 0.f;
  for (int32_t k = beg; k < end; ++k) l += __expf(leaky(as[(uint32_t)col[k] * (uint32_t)H] + adi, slope) - m);
  const float rl = 1.f / (l + 1e-16f);
  float* ac = alpha_csr + b * nnz * H + h;
  float* ap = alpha_pyg ? alpha_pyg + b * nnz * H + h : nullptr;
  for (int32_t k = beg; k < end; ++k) {
    const float al = __expf(leaky(as[(uint32_t)col[k] * (uint32_t)H] + adi, slope) - m) * rl;
    ac[(int64_t)k * H] = al;
    if (ap) ap[(int64_t)perm[k] * H] = al;
  }
}

// Butterfly transpose-reduce: every lane of an L-lane group holds L partial values p[0..L); afterwards lane g
// holds the sum over the group's lanes of p[g].  L - 1 shuffles for L reductions (a shuffle tree per value would
// be L log2 L).
template <int L>
__device__ __forceinline__ float xreduce(float (&p)[L], int gl, unsigned mask) {
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
__device__ __forceinline__ float gsum_strided(float v, unsigned mask) {
#pragma unroll
  for (int o = SB; o < L; o <<= 1) v += __shfl_xor_sync(mask, v, o, L);
  return v;
}

// Backward pass 1, group per (receiver i, SB samples):
//   dalpha_k = <do_h(i), z[col_k, h]>;  t = sum_k alpha_k dalpha_k;  g_k = alpha_k (dalpha_k - t) * LeakyReLU'
//   g_csr[b,k,h] = g_k;  da_dst[b,i,h] = sum_k g_k
// The row is walked in batches of NB = L / SB neighbours: the L partial dot products (NB neighbours x SB
// samples) of a batch are transpose-reduced at once, which leaves lane l with the finished dalpha of
// (neighbour l / SB, sample l % SB); everything after that is per-lane scalar work on its own (edge, sample).
template <int VW, int L, int SB>
__global__ void __launch_bounds__(kWarps * 32, SB >= 4 ? 2 : 4)
    gat_bwd_dst_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                       const float* __restrict__ z, const float* __restrict__ a_src,
                       const float* __restrict__ a_dst, const float* __restrict__ alpha_csr,
                       const float* __restrict__ dout, float* __restrict__ g_csr, float* __restrict__ da_dst,
                       int64_t N, int64_t nnz, int B, int H, int C, int concat, float slope) {
  using V = W<VW>;
  constexpr int kGroups = 32 / L, NB = L / SB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane & (L - 1);
  const unsigned mask = group_mask<L>(lane);
  const int64_t i = ((int64_t)blockIdx.x * kWarps + warp) * kGroups + lane / L;
  if (i >= N) return;
  const int b0 = blockIdx.y * SB;
  const int nb = min(SB, B - b0);
  const int32_t beg = rowptr[i], end = rowptr[i + 1];
  const int HC = H * C, Cout = concat ? HC : C;
  const int words = (C + VW - 1) / VW, nchunks = (words + L - 1) / L;
  const float hs = concat ? 1.f : 1.f / (float)H;
  const bool single = (end - beg) <= L;    // at most SB batches: dalpha stays in registers
  const bool one = nchunks == 1;           // the usual case: this lane's slice of do(i) is loaded once per head
  const int jj = gl / SB, s_me = gl % SB;  // after the transpose-reduce this lane owns (neighbour jj, sample s_me)
  const bool s_ok = s_me < nb;
  const int64_t bs = b0 + s_me;

  for (int h = 0; h < H; ++h) {
    const float* dbase = dout + (concat ? h * C : 0);
    const int gofs = gl * VW < C ? gl * VW : 0;             // idle lanes re-read word 0 (their products are dropped)
    const bool glive = gl * VW < C;
    typename V::T dv[SB];
    const float* zp[SB];                                    // z[b, 0, h, gofs], ragged sample blocks re-read nb - 1
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      const int64_t sb = b0 + min(s, nb - 1);
      zp[s] = z + sb * N * HC + h * C + gofs;
      dv[s] = (one && glive) ? V::load(dbase + (sb * N + i) * Cout + gofs) : V::zero();
    }
    // finished dalpha (scaled by the head-mean factor) of this lane's (edge, sample) in the batch starting at `base`
    auto batch = [&](int32_t base) -> float {
      float p[L];
      if (one) {
        typename V::T zv[NB][SB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int32_t k = min(base + j, end - 1);          // past the row end: re-read the last edge, dropped below
          const uint32_t ro = (uint32_t)col[k] * (uint32_t)HC;
#pragma unroll
          for (int s = 0; s < SB; ++s) zv[j][s] = V::load(zp[s] + ro);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j)
#pragma unroll
          for (int s = 0; s < SB; ++s) p[j * SB + s] = V::dot(dv[s], zv[j][s]);
      } else {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int32_t k = min(base + j, end - 1);
          const uint32_t ro = (uint32_t)col[k] * (uint32_t)HC;
#pragma unroll
          for (int s = 0; s < SB; ++s) p[j * SB + s] = 0.f;
          for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int off = (chunk * L + gl) * VW;
            if (off < C) {
#pragma unroll
              for (int s = 0; s < SB; ++s) {
                const int64_t sb = b0 + min(s, nb - 1);
                p[j * SB + s] += V::dot(V::load(dbase + (sb * N + i) * Cout + off), V::load(zp[s] - gofs + off + ro));
              }
            }
          }
        }
      }
      return xreduce<L>(p, gl, mask) * hs;
    };

    const float adst = s_ok ? a_dst[(bs * N + i) * H + h] : 0.f;
    float tp = 0.f, gs = 0.f;
    if (single) {
      float dal[SB], al[SB];
#pragma unroll
      for (int q = 0; q < SB; ++q) {
        dal[q] = al[q] = 0.f;
        const int32_t base = beg + q * NB;
        if (base < end) {                                   // group-uniform
          const float r = batch(base);
          const int32_t k = base + jj;
          if (k < end && s_ok) {
            dal[q] = r;
            al[q] = alpha_csr[(bs * nnz + k) * H + h];
            tp += al[q] * r;
          }
        }
      }
      const float t = gsum_strided<L, SB>(tp, mask);
#pragma unroll
      for (int q = 0; q < SB; ++q) {
        const int32_t k = beg + q * NB + jj;
        if (k < end && s_ok) {
          const float pre = a_src[(bs * N + col[k]) * H + h] + adst;
          const float g = al[q] * (dal[q] - t) * (pre > 0.f ? 1.f : slope);
          g_csr[(bs * nnz + k) * H + h] = g;
          gs += g;
        }
      }
    } else {
      for (int32_t base = beg; base < end; base += NB) {
        const float r = batch(base);
        const int32_t k = base + jj;
        if (k < end && s_ok) {
          const int64_t idx = (bs * nnz + k) * H + h;
          tp += alpha_csr[idx] * r;
          g_csr[idx] = r;                                  // staged; re-read below by the same lane
        }
      }
      const float t = gsum_strided<L, SB>(tp, mask);
      for (int32_t base = beg; base < end; base += NB) {
        const int32_t k = base + jj;
        if (k < end && s_ok) {
          const int64_t idx = (bs * nnz + k) * H + h;
          const float pre = a_src[(bs * N + col[k]) * H + h] + adst;
          const float g = alpha_csr[idx] * (g_csr[idx] - t) * (pre > 0.f ? 1.f : slope);
          g_csr[idx] = g;
          gs += g;
        }
      }
    }
    gs = gsum_strided<L, SB>(gs, mask);
    if (jj == 0 && s_ok) da_dst[(bs * N + i) * H + h] = gs;
  }
}

// Backward pass 2, group per (sender j, SB samples), sender-grouped CSR:
//   da_src[b,j,h] = sum_k g_k ;  dz[b,j,h,:] = sum_k alpha_k do_h(i_k) + da_src att_src[h] + da_dst att_dst[h]
template <int VW, int L, int SB>
__global__ void __launch_bounds__(kWarps * 32, SB >= 4 ? 2 : 4)
    gat_bwd_src_kernel(const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ col_t,
                       const int32_t* __restrict__ t2r, const float* __restrict__ alpha_csr,
                       const float* __restrict__ g_csr, const float* __restrict__ att_src,
                       const float* __restrict__ att_dst, const float* __restrict__ dout,
                       const float* __restrict__ da_dst, float* __restrict__ da_src, float* __restrict__ dz,
                       int64_t N, int64_t nnz, int B, int H, int C, int concat) {
  using V = W<VW>;
  constexpr int kGroups = 32 / L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane & (L - 1);
  const unsigned mask = group_mask<L>(lane);
  const int64_t jn = ((int64_t)blockIdx.x * kWarps + warp) * kGroups + lane / L;
  if (jn >= N) return;
  const int b0 = blockIdx.y * SB;
  const int nb = min(SB, B - b0);
  const int32_t beg = rowptr_t[jn], end = rowptr_t[jn + 1];
  const int HC = H * C, Cout = concat ? HC : C;
  const int words = (C + VW - 1) / VW, nchunks = (words + L - 1) / L;
  const float hs = concat ? 1.f : 1.f / (float)H;
  int64_t sb_[SB];                                          // (clamped) sample index
#pragma unroll
  for (int s = 0; s < SB; ++s) sb_[s] = b0 + min(s, nb - 1);

  for (int h = 0; h < H; ++h) {
    const float* gp[SB];                                    // g_csr[b, 0, h], alpha_csr[b, 0, h]
    const float* ap[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      gp[s] = g_csr + sb_[s] * nnz * H + h;
      ap[s] = alpha_csr + sb_[s] * nnz * H + h;
    }
    float gsv[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) gsv[s] = 0.f;
    for (int32_t k = beg + gl; k < end; k += L) {
      const uint32_t kr = (uint32_t)t2r[k] * (uint32_t)H;
#pragma unroll
      for (int s = 0; s < SB; ++s) gsv[s] += gp[s][kr];
    }
    float dad[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      gsv[s] = gsum<L>(gsv[s], mask);
      dad[s] = da_dst[(sb_[s] * N + jn) * H + h];
      if (gl == 0 && s < nb) da_src[(sb_[s] * N + jn) * H + h] = gsv[s];
    }
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int off = (chunk * L + gl) * VW;
      const bool live = off < C;
      const int offc = live ? off : 0;
      const float* dp[SB];                                  // dout[b, 0, (h), offc]
#pragma unroll
      for (int s = 0; s < SB; ++s) dp[s] = dout + sb_[s] * N * Cout + (concat ? h * C : 0) + offc;
      typename V::T acc[SB];
#pragma unroll
      for (int s = 0; s < SB; ++s) acc[s] = V::zero();
      for (int32_t base = beg; base < end; base += L) {
        const int32_t k = base + gl;
        const int n = min(L, end - base);
        int32_t i_reg = 0;
        float a_reg[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s) a_reg[s] = 0.f;
        if (k < end) {
          i_reg = col_t[k];
          const uint32_t kr = (uint32_t)t2r[k] * (uint32_t)H;
#pragma unroll
          for (int s = 0; s < SB; ++s) a_reg[s] = ap[s][kr] * hs;
        }
#pragma unroll kGatUnroll
        for (int e = 0; e < n; ++e) {
          const uint32_t ro = (uint32_t)__shfl_sync(mask, i_reg, e, L) * (uint32_t)Cout;
          float a[SB];
          typename V::T v[SB];
#pragma unroll
          for (int s = 0; s < SB; ++s) {
            a[s] = __shfl_sync(mask, a_reg[s], e, L);
            v[s] = V::load(dp[s] + ro);
          }
#pragma unroll
          for (int s = 0; s < SB; ++s) V::fma(acc[s], a[s], v[s]);
        }
      }
      if (live) {
        const typename V::T as = V::load(att_src + h * C + off), ad = V::load(att_dst + h * C + off);
#pragma unroll
        for (int s = 0; s < SB; ++s) {
          if (s < nb) {
            V::fma(acc[s], gsv[s], as);
            V::fma(acc[s], dad[s], ad);
            V::store(dz + ((int64_t)(b0 + s) * N + jn) * HC + h * C + off, acc[s]);
          }
        }
      }
    }
  }
}

// part[blk][0][hc] = sum_rows da_src[row,h] z[row,hc];  part[blk][1][hc] likewise with da_dst
__global__ void gat_datt_partial_kernel(const float* __restrict__ z, const float* __restrict__ da_src,
                                        const float* __restrict__ da_dst, float* __restrict__ part, int64_t rows,
                                        int H, int C, int64_t rows_per_block) {
  __shared__ float sm[2][8][33];
  const int HC = H * C;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int c0 = 0; c0 < HC; c0 += 32) {
    const int c = c0 + threadIdx.x;
    float s = 0.f, d = 0.f;
    if (c < HC) {
      const int h = c / C;
      for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
        const float v = z[r * HC + c];
        s = fmaf(da_src[r * H + h], v, s);
        d = fmaf(da_dst[r * H + h], v, d);
      }
    }
    sm[0][threadIdx.y][threadIdx.x] = s;
    sm[1][threadIdx.y][threadIdx.x] = d;
    __syncthreads();
    if (threadIdx.y < 2 && c < HC) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[threadIdx.y][k][threadIdx.x];
      part[((int64_t)blockIdx.x * 2 + threadIdx.y) * HC + c] = t;
    }
    __syncthreads();
  }
}

// 128-bit version (C % 4 == 0, H C <= 1024, 16-byte aligned z): float4 column groups x row lanes, 4 rows in flight.
__global__ void __launch_bounds__(256)
    gat_datt_partial_v4_kernel(const float4* __restrict__ z, const float* __restrict__ da_src,
                               const float* __restrict__ da_dst, float* __restrict__ part, int64_t rows, int H, int C,
                               int64_t rows_per_block) {
  __shared__ float4 sm[2][256];
  const int HC = H * C, ncol = HC >> 2, lanes = 256 / ncol, tid = threadIdx.x;
  const int ci = tid % ncol, rl = tid / ncol, h = (4 * ci) / C;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), d = s;
  if (rl < lanes) {
    for (int64_t r = r0 + rl; r < r1; r += 4 * lanes) {
      float4 v[4];
      float as[4], ad[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t rr = r + (int64_t)u * lanes;
        const bool ok = rr < r1;
        v[u] = ok ? __ldg(z + rr * ncol + ci) : make_float4(0.f, 0.f, 0.f, 0.f);
        as[u] = ok ? __ldg(da_src + rr * H + h) : 0.f;
        ad[u] = ok ? __ldg(da_dst + rr * H + h) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s.x = fmaf(as[u], v[u].x, s.x); s.y = fmaf(as[u], v[u].y, s.y);
        s.z = fmaf(as[u], v[u].z, s.z); s.w = fmaf(as[u], v[u].w, s.w);
        d.x = fmaf(ad[u], v[u].x, d.x); d.y = fmaf(ad[u], v[u].y, d.y);
        d.z = fmaf(ad[u], v[u].z, d.z); d.w = fmaf(ad[u], v[u].w, d.w);
      }
    }
  }
  sm[0][tid] = s;
  sm[1][tid] = d;
  __syncthreads();
  if (tid < 2 * ncol) {
    const int which = tid / ncol, cc = tid % ncol;
    float4 t = sm[which][cc];
    for (int l = 1; l < lanes; ++l) {
      const float4 o = sm[which][l * ncol + cc];
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    float* dst = part + ((int64_t)blockIdx.x * 2 + which) * HC + 4 * cc;
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
  }
}

// out[i] = sum_k part[k * 2n + i], fixed order (32 interleaved partial sums, then ascending); block (32, 32)
__global__ void reduce2_kernel(const float* __restrict__ part, int nblk, int n, float* __restrict__ out0,
                               float* __restrict__ out1) {
  __shared__ float sm[32][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (i < 2 * n)
#pragma unroll 4
    for (int k = threadIdx.y; k < nblk; k += 32) s += part[(int64_t)k * 2 * n + i];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < 2 * n) {
    float t = sm[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < 32; ++y) t += sm[y][threadIdx.x];
    if (i < n) out0[i] = t;
    else out1[i - n] = t;
  }
}

__global__ void prune_flags_kernel(const float* __restrict__ alpha, int64_t nnz, float thr,
                                   int32_t* __restrict__ flag) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < nnz) flag[p] = alpha[p] >= thr;
}

__global__ void prune_compact_kernel(const int64_t* __restrict__ ei, int64_t ei_stride,
                                     const int32_t* __restrict__ flag, const int32_t* __restrict__ pos,
                                     int64_t nnz, int64_t* __restrict__ out, int64_t out_stride,
                                     int32_t* __restrict__ count) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p == 0) *count = pos[nnz];
  if (p < nnz && flag[p]) {
    out[pos[p]] = ei[p];
    out[out_stride + pos[p]] = ei[ei_stride + p];
  }
}

struct Plan {
  int nblk;
  int64_t rows_per_block;
};
Plan rows_plan(int64_t rows) {
  int64_t nblk = 4 * kNumSMs;
  int64_t rpb = ceil_div(rows, nblk);
  if (rpb < 8) rpb = 8;
  nblk = ceil_div(rows, rpb);
  if (nblk < 1) nblk = 1;
  return {(int)nblk, rpb};
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int pick_l(int words) { return words <= 4 ? 4 : words <= 8 ? 8 : words <= 16 ? 16 : 32; }

}  // namespace
}  // namespace gcl

namespace gcl {
int spmm_run(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, float* out, int64_t batch,
             int64_t n_rows_out, int64_t n_rows_in, int64_t channels, int64_t x_bstride, int64_t out_bstride,
             const float* bias, const float* prelu_slope, float* z_out, int64_t nnz, int64_t w_bstride, float w_scale,
             cudaStream_t s);
}
using namespace gcl;

#define GAT_DISPATCH_L(VWC, SBC, LV, KERNEL, ...)                                        \
  do {                                                                                    \
    if (LV == 4) KERNEL<VWC, 4, SBC><<<grid, kWarps * 32, 0, s>>>(__VA_ARGS__);           \
    else if (LV == 8) KERNEL<VWC, 8, SBC><<<grid, kWarps * 32, 0, s>>>(__VA_ARGS__);      \
    else if (LV == 16) KERNEL<VWC, 16, SBC><<<grid, kWarps * 32, 0, s>>>(__VA_ARGS__);    \
    else KERNEL<VWC, 32, SBC><<<grid, kWarps * 32, 0, s>>>(__VA_ARGS__);                  \
  } while (0)
#define GAT_DISPATCH(VWV, SBV, LV, KERNEL, ...)                                            \
  do {                                                                                    \
    if (VWV == 4) {                                                                       \
      if (SBV == 4) GAT_DISPATCH_L(4, 4, LV, KERNEL, __VA_ARGS__);                        \
      else if (SBV == 2) GAT_DISPATCH_L(4, 2, LV, KERNEL, __VA_ARGS__);                   \
      else GAT_DISPATCH_L(4, 1, LV, KERNEL, __VA_ARGS__);                                 \
    } else {                                                                              \
      GAT_DISPATCH_L(1, 1, LV, KERNEL, __VA_ARGS__);                                      \
    }                                                                                     \
  } while (0)

// samples per lane group (the scalar VW = 1 fallback keeps one)
static int gat_pick_sb(int vw, int64_t B, int64_t N, int64_t HC) {
  if (vw != 4) return 1;
  static const int cap = getenv("GCL_GAT_SB") ? atoi(getenv("GCL_GAT_SB")) : 4;
  int sb = cap;
  while (sb > 1 && (sb > B || sb * N * HC * 4 > (48ll << 20))) sb >>= 1;
  return sb;
}

extern "C" int gcl_gat_scores_f32(const float* z, const float* att_src, const float* att_dst, float* a_src,
                                  float* a_dst, int64_t rows, int64_t heads, int64_t c, void* stream) {
  GCL_CHECK_ARG(z && att_src && att_dst && a_src && a_dst && rows >= 0 && heads > 0 && c > 0,
                "gcl_gat_scores_f32: bad argument");
  if (rows == 0) return GCL_OK;
  gat_scores_kernel<<<(unsigned)ceil_div(rows * heads * 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, att_src, att_dst, a_src, a_dst, rows, (int)heads, (int)c);
  GCL_CHECK_LAUNCH("gcl_gat_scores_f32");
  return GCL_OK;
}

extern "C" int gcl_gat_fwd_f32(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const float* z,
                               const float* a_src, const float* a_dst, const float* bias, float* out,
                               float* alpha_csr, float* alpha_pyg, const float* prelu_slope, float* z_out,
                               int64_t batch, int64_t n_nodes, int64_t nnz,
                               int64_t heads, int64_t c, int concat, float negative_slope, void* stream) {
  GCL_CHECK_ARG(rowptr && col && z && a_src && a_dst && out && alpha_csr, "gcl_gat_fwd_f32: null pointer argument");
  GCL_CHECK_ARG(!alpha_pyg || perm, "gcl_gat_fwd_f32: alpha_pyg needs perm");
  GCL_CHECK_ARG(n_nodes * heads * c < (1ll << 31) && nnz * heads < (1ll << 31),
                "gcl_gat_fwd_f32: one sample's features / attention entries must index with 31 bits");
  GCL_CHECK_ARG(batch >= 0 && batch <= 65535 && n_nodes >= 0 && nnz >= 0 && heads > 0 && c > 0,
                "gcl_gat_fwd_f32: bad sizes");
  if (batch == 0 || n_nodes == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const bool fused_only = getenv("GCL_GAT_FUSED") && getenv("GCL_GAT_FUSED")[0] == '1';
  if (heads == 1 && !fused_only) {
    // single head: coefficients by gat_alpha_kernel, aggregation = SpMM with per-sample weights
    gat_alpha_kernel<<<(unsigned)ceil_div(batch * n_nodes, 256), 256, 0, s>>>(rowptr, col, perm, a_src, a_dst, alpha_csr,
                                                                              alpha_pyg, n_nodes, nnz, (int)batch, 1,
                                                                              negative_slope);
    GCL_CHECK_LAUNCH("gcl_gat_fwd_f32(alpha)");
    return spmm_run(rowptr, col, alpha_csr, z, out, batch, n_nodes, n_nodes, c, n_nodes * c, n_nodes * c, bias,
                    prelu_slope, z_out, nnz, nnz, 1.f, s);
  }
  if (prelu_slope) {
    set_error("gcl_gat_fwd_f32: the fused PReLU epilogue exists for heads == 1 only");
    return GCL_ERR_UNSUPPORTED;
  }
  const int vw = (c % 4 == 0 && al16(z) && al16(out) && (!bias || al16(bias))) ? 4 : 1;
  const int l = pick_l((int)ceil_div(c, vw));
  const int sb = gat_pick_sb(vw, batch, n_nodes, heads * c);
  dim3 grid((unsigned)ceil_div(n_nodes, (int64_t)kWarps * (32 / l)), (unsigned)ceil_div(batch, sb));
  GAT_DISPATCH(vw, sb, l, gat_fwd_kernel, rowptr, col, perm, z, a_src, a_dst, bias, out, alpha_csr, alpha_pyg, n_nodes,
               nnz, (int)batch, (int)heads, (int)c, concat, negative_slope);
  GCL_CHECK_LAUNCH("gcl_gat_fwd_f32");
  return GCL_OK;
}

extern "C" int gcl_gat_bwd_f32(const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                               const int32_t* col_t, const int32_t* t2r, const float* z, const float* a_src,
                               const float* a_dst, const float* alpha_csr, const float* att_src,
                               const float* att_dst, const float* dout, float* g_csr, float* da_src, float* da_dst,
                               float* dz, int64_t batch, int64_t n_nodes, int64_t nnz, int64_t heads, int64_t c,
                               int concat, float negative_slope, void* stream) {
  GCL_CHECK_ARG(rowptr && col && rowptr_t && col_t && t2r && z && a_src && a_dst && alpha_csr && att_src && att_dst &&
                    dout && g_csr && da_src && da_dst && dz,
                "gcl_gat_bwd_f32: null pointer argument");
  GCL_CHECK_ARG(n_nodes * heads * c < (1ll << 31) && nnz * heads < (1ll << 31),
                "gcl_gat_bwd_f32: one sample's features / attention entries must index with 31 bits");
  GCL_CHECK_ARG(batch >= 0 && batch <= 65535 && n_nodes >= 0 && nnz >= 0 && heads > 0 && c > 0,
                "gcl_gat_bwd_f32: bad sizes");
  if (batch == 0 || n_nodes == 0) return GCL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vw = (c % 4 == 0 && al16(z) && al16(dout) && al16(dz) && al16(att_src) && al16(att_dst)) ? 4 : 1;
  const int l = pick_l((int)ceil_div(c, vw));
  const int sb = gat_pick_sb(vw, batch, n_nodes, heads * c);
  dim3 grid((unsigned)ceil_div(n_nodes, (int64_t)kWarps * (32 / l)), (unsigned)ceil_div(batch, sb));
  GAT_DISPATCH(vw, sb, l, gat_bwd_dst_kernel, rowptr, col, z, a_src, a_dst, alpha_csr, dout, g_csr, da_dst, n_nodes,
               nnz, (int)batch, (int)heads, (int)c, concat, negative_slope);
  GCL_CHECK_LAUNCH("gcl_gat_bwd_f32(dst pass)");
  GAT_DISPATCH(vw, sb, l, gat_bwd_src_kernel, rowptr_t, col_t, t2r, alpha_csr, g_csr, att_src, att_dst, dout, da_dst,
               da_src, dz, n_nodes, nnz, (int)batch, (int)heads, (int)c, concat);
  GCL_CHECK_LAUNCH("gcl_gat_bwd_f32(src pass)");
  return GCL_OK;
}

extern "C" size_t gcl_gat_datt_workspace_bytes(int64_t rows, int64_t heads, int64_t c) {
  if (rows < 0 || heads <= 0 || c <= 0) return 0;
  return (size_t)rows_plan(rows).nblk * 2 * (size_t)(heads * c) * sizeof(float) + 256;
}

extern "C" int gcl_gat_datt_f32(const float* z, const float* da_src, const float* da_dst, float* datt_src,
                                float* datt_dst, int64_t rows, int64_t heads, int64_t c, void* workspace,
                                size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(z && da_src && da_dst && datt_src && datt_dst && workspace && rows >= 0 && heads > 0 && c > 0,
                "gcl_gat_datt_f32: bad argument");
  if (workspace_bytes < gcl_gat_datt_workspace_bytes(rows, heads, c)) {
    set_error("gcl_gat_datt_f32: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int HC = (int)(heads * c);
  if (rows == 0) {
    cudaMemsetAsync(datt_src, 0, sizeof(float) * HC, s);
    cudaMemsetAsync(datt_dst, 0, sizeof(float) * HC, s);
    return GCL_OK;
  }
  Plan pl = rows_plan(rows);
  float* part = static_cast<float*>(workspace);
  if ((c & 3) == 0 && HC <= 1024 && 2 * (HC >> 2) <= 256 && (reinterpret_cast<uintptr_t>(z) & 15u) == 0)
    gat_datt_partial_v4_kernel<<<pl.nblk, 256, 0, s>>>(reinterpret_cast<const float4*>(z), da_src, da_dst, part, rows,
                                                       (int)heads, (int)c, pl.rows_per_block);
  else
    gat_datt_partial_kernel<<<pl.nblk, dim3(32, 8), 0, s>>>(z, da_src, da_dst, part, rows, (int)heads, (int)c,
                                                            pl.rows_per_block);
  GCL_CHECK_LAUNCH("gcl_gat_datt_f32(partial)");
  reduce2_kernel<<<(unsigned)ceil_div(2 * HC, 32), dim3(32, 32), 0, s>>>(part, pl.nblk, HC, datt_src, datt_dst);
  GCL_CHECK_LAUNCH("gcl_gat_datt_f32(reduce)");
  return GCL_OK;
}

extern "C" size_t gcl_edge_prune_workspace_bytes(int64_t nnz) {
  if (nnz < 0) return 0;
  return 2 * (((size_t)(nnz + 1) * sizeof(int32_t) + 255) & ~size_t(255)) + 256;
}

extern "C" int gcl_edge_prune(const int64_t* ei_pyg, const float* alpha_pyg, int64_t nnz, int64_t ei_stride,
                              float threshold, int64_t* ei_kept, int64_t kept_stride, int32_t* count_out,
                              void* workspace, size_t workspace_bytes, void* stream) {
  GCL_CHECK_ARG(ei_pyg && alpha_pyg && ei_kept && count_out && workspace && nnz >= 0, "gcl_edge_prune: bad argument");
  if (workspace_bytes < gcl_edge_prune_workspace_bytes(nnz)) {
    set_error("gcl_edge_prune: workspace too small");
    return GCL_ERR_WORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int32_t* flag = static_cast<int32_t*>(workspace);
  int32_t* pos = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) +
                                            (((size_t)(nnz + 1) * sizeof(int32_t) + 255) & ~size_t(255)));
  if (nnz > 0) {
    prune_flags_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, s>>>(alpha_pyg, nnz, threshold, flag);
    GCL_CHECK_LAUNCH("gcl_edge_prune(flags)");
  }
  scan_exclusive_kernel<<<1, kScanThreads, 0, s>>>(flag, pos, nnz, nullptr);
  GCL_CHECK_LAUNCH("gcl_edge_prune(scan)");
  prune_compact_kernel<<<(unsigned)ceil_div(nnz + 1, 256), 256, 0, s>>>(ei_pyg, ei_stride, flag, pos, nnz, ei_kept,
                                                                        kept_stride, count_out);
  GCL_CHECK_LAUNCH("gcl_edge_prune(compact)");
  return GCL_OK;
}

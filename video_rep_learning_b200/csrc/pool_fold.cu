// Rank-E folded entity pooling: the cross-attention of LSTPCrossAtt (mvformer.py:352-414) has only E (3..16) static
// queries and one head, so K = X Wk^T + bk and V = X Wv^T + bv never need to exist:
//
//   scores[e,p] = (Q_s[e]+Q_b) . K[p] / sqrt(SPC) = Wq[e,:] . X[p,:] + const(e)      Wq = (Q_s+Q_b) Wk / sqrt(SPC)  [E, C_in]
//   A[e,:]      = softmax_p(scores[e,:])                                             (the constant cancels)
//   ent[e,:]    = sum_p A[e,p] V[p,:] = (sum_p A[e,p] X[p,:]) Wv^T + bv = px[e,:] Wv^T + bv       (rows of A sum to 1)
//
// which turns 2*2*P*C_in*SPC FLOPs per frame (96 % of the whole step at the BASELINE shapes) into ONE streaming pass
// over the patch tokens: 4*E*P*C_in FLOPs on CUDA cores, bound by the HBM read of X.  Backward is the same kind of pass:
//
//   G[e,:]   = dEnt[e,:] Wv                     (small GEMM)          delta[e] = G[e,:] . px[e,:] = sum_p A[e,p] dA[e,p]
//   dA[e,p]  = G[e,:] . X[p,:]                                        dS[e,p]  = A[e,p] (dA[e,p] - delta[e])
//   dWq[e,:] = sum_{frames,p} dS[e,p] X[p,:]    -> dWk = Q^T dWq / sqrt(SPC), dQ = dWq Wk^T / sqrt(SPC);  dWv = dEnt^T px
//
// Kernel shape (both directions): persistent CTAs, one frame at a time; the token rows of a frame are contiguous in
// the ViT-native token-major layout, so groups of 8 tokens (8*C_in*sizeof(T) bytes) are streamed into a shared-memory
// ring by bulk async copies (cp.async.bulk + mbarrier complete_tx).  Thread t owns channels [8t, 8t+8) for the whole
// kernel: its slice of Wq (or G) and of the running accumulators lives in registers, every 16-byte shared-memory read
// is conflict-free, and the only cross-thread traffic is the reduction of the 8*E partial dot products of a group
// (recursive-halving warp shuffles, then one shared-memory hop across warps).  Forward keeps an online softmax
// (running max / sum per entity, accumulators rescaled when the max moves), so X is read exactly once.
#include <math.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace mvf {
namespace fold {

constexpr int TG = 8;         // tokens per group (one pipeline stage)
constexpr int MAX_STAGES = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {  // a protocol bug must trap, never hang the GPU
      printf("mvf pool_fold: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- packed fp32 pairs: Blackwell issues fma.rn.f32x2 (FFMA2) at the rate of a scalar FFMA, and the FMA pipe is what
// bounds these kernels (3 warps per scheduler x ~400 FFMA per 8-token group otherwise) ---------------------------------
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float a, float b) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 up2(f2_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f2_t ffma2(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2_t fmul2(f2_t a, f2_t b) {
  f2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---- the 8 channels a thread owns in one token row, as 4 fp32 pairs ------------------------------------------------------
template <typename T> struct XV;
template <> struct XV<bf16> {
  static __device__ __forceinline__ void load(const uint8_t* row, int cv, f2_t* x2) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + cv * 16);
    x2[0] = pk2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u));
    x2[1] = pk2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    x2[2] = pk2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u));
    x2[3] = pk2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
  }
};
template <> struct XV<float> {
  static __device__ __forceinline__ void load(const uint8_t* row, int cv, f2_t* x2) {
    const float4 a = *reinterpret_cast<const float4*>(row + cv * 32);
    const float4 b = *reinterpret_cast<const float4*>(row + cv * 32 + 16);
    x2[0] = pk2(a.x, a.y); x2[1] = pk2(a.z, a.w); x2[2] = pk2(b.x, b.y); x2[3] = pk2(b.z, b.w);
  }
};

// x2[p][k]: token p of the group, channel pair k of this thread (zero for tokens past the end of the frame)
template <typename T>
__device__ __forceinline__ void load_group(const uint8_t* st, int row_bytes, int cv, bool act, int ntok, f2_t (&x2)[8][4]) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    if (act && p < ntok) XV<T>::load(st + (size_t)p * row_bytes, cv, x2[p]);
    else x2[p][0] = x2[p][1] = x2[p][2] = x2[p][3] = 0ull;
  }
}
// part[e*8+p] = sum over this thread's 8 channels of v[e][c] * x[p][c]
template <int E>
__device__ __forceinline__ void dot_group(const f2_t (&v2)[E][4], const f2_t (&x2)[8][4], float* part) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      f2_t s2 = fmul2(v2[e][0], x2[p][0]);
#pragma unroll
      for (int k = 1; k < 4; ++k) s2 = ffma2(v2[e][k], x2[p][k], s2);
      const float2 s = up2(s2);
      part[e * 8 + p] = s.x + s.y;
    }
  }
}
// acc[e][c] += w[e][p] * x[p][c]   (w: E*8 floats in shared memory, broadcast reads)
template <int E>
__device__ __forceinline__ void axpy_group(const float* wb, const f2_t (&x2)[8][4], f2_t (&acc2)[E][4]) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const float4 w0 = *reinterpret_cast<const float4*>(wb + e * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(wb + e * 8 + 4);
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const f2_t ww = pk2(w[p], w[p]);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc2[e][k] = ffma2(ww, x2[p][k], acc2[e][k]);
    }
  }
}
template <int E>
__device__ __forceinline__ void load_vec(const float* src, bool act, f2_t (&v2)[E][4], size_t stride) {
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (act) {
      const float4 a = *reinterpret_cast<const float4*>(src + e * stride);
      const float4 b = *reinterpret_cast<const float4*>(src + e * stride + 4);
      v2[e][0] = pk2(a.x, a.y); v2[e][1] = pk2(a.z, a.w); v2[e][2] = pk2(b.x, b.y); v2[e][3] = pk2(b.z, b.w);
    } else {
      v2[e][0] = v2[e][1] = v2[e][2] = v2[e][3] = 0ull;
    }
  }
}

// ---- recursive-halving warp reduction of N per-lane values: lane L ends with the full sum of value red_index(L) ------
// (N + ~5 shuffles instead of 5 N).  Odd sizes are padded with a zero; lanes that end on padding get index -1.
template <int N, int BIT>
struct Red {
  static __device__ __forceinline__ void run(float* v, int lane) {
    if constexpr (BIT > 0) {
      if constexpr (N == 1) {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], BIT);
        Red<1, BIT / 2>::run(v, lane);
      } else {
        constexpr int NP = (N + 1) & ~1, H = NP / 2;
        if constexpr (NP != N) v[N] = 0.f;
        const bool up = (lane & BIT) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const float send = up ? v[i] : v[i + H];
          const float keep = up ? v[i + H] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
        }
        Red<H, BIT / 2>::run(v, lane);
      }
    }
  }
  static __device__ __forceinline__ int index(int lane) {
    if constexpr (BIT == 0) return 0;
    else if constexpr (N == 1) return 0;
    else {
      constexpr int NP = (N + 1) & ~1, H = NP / 2;
      const int sub = Red<H, BIT / 2>::index(lane);
      const int pos = ((lane & BIT) ? H : 0) + sub;
      return (sub < 0 || pos >= N) ? -1 : pos;
    }
  }
};

struct Geom {
  int F, P, C;        // frames, tokens per frame, channels
  int Etot, e0;       // entities in the model, first entity of this pass
  int stages;
  int row_bytes;      // C * sizeof(T)
};

// shared-memory carve-up (bytes), identical for both directions
struct Carve {
  size_t ring, bars, partial, wbuf, table, misc, total;
};
static Carve carve(int stages, int row_bytes, int nwarps, int E, int P) {
  Carve c;
  c.ring = (size_t)stages * TG * row_bytes;
  c.bars = 8 * MAX_STAGES;
  c.partial = (size_t)2 * nwarps * 32 * 4;
  c.wbuf = (size_t)nwarps * 40 * 4;
  c.table = ((size_t)E * P * 4 + 15) / 16 * 16;   // raw scores (fwd) / probabilities (bwd) of the current frame
  c.misc = ((size_t)(2 * E + nwarps * E) * 4 + 15) / 16 * 16;
  c.total = c.ring + c.bars + c.partial + c.wbuf + c.table + c.misc;
  return c;
}

#define MVF_FOLD_SMEM_SETUP()                                                                         \
  extern __shared__ __align__(128) uint8_t smraw[];                                                   \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = blockDim.x >> 5;                 \
  uint8_t* ring = smraw;                                                                              \
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)g.stages * TG * g.row_bytes);           \
  float* partial = reinterpret_cast<float*>(full + MAX_STAGES);                                       \
  float* wbuf = partial + 2 * NW * 32;                                                                \
  float* table = wbuf + NW * 40;                                                                      \
  float* misc = table + (((size_t)E * g.P + 3) / 4) * 4;                                              \
  const int cv = tid;                                                                                 \
  const bool act = cv < (g.C >> 3);                                                                   \
  const int nG = (g.P + TG - 1) / TG;                                                                 \
  const int nF = blockIdx.x < g.F ? (g.F - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;               \
  const int total = nF * nG;                                                                          \
  if (tid == 0) {                                                                                     \
    for (int s = 0; s < g.stages; ++s) mbar_init(&full[s], 1);                                        \
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                                \
  }                                                                                                   \
  __syncthreads();                                                                                    \
  auto issue = [&](int n2) {                                                                          \
    const int fi2 = n2 / nG, g2 = n2 - fi2 * nG;                                                      \
    const int64_t f2 = blockIdx.x + (int64_t)fi2 * gridDim.x;                                         \
    const int nt2 = min(TG, g.P - g2 * TG);                                                           \
    const uint32_t bytes = (uint32_t)nt2 * (uint32_t)g.row_bytes;                                     \
    const int s2 = n2 % g.stages;                                                                     \
    mbar_expect_tx(&full[s2], bytes);                                                                 \
    bulk_g2s(ring + (size_t)s2 * TG * g.row_bytes,                                                    \
             reinterpret_cast<const uint8_t*>(X) + (f2 * g.P + (int64_t)g2 * TG) * g.row_bytes, bytes, &full[s2]); \
  };                                                                                                  \
  if (tid == 0)                                                                                       \
    for (int n2 = 0; n2 < g.stages && n2 < total; ++n2) issue(n2);

// =====================================================================================================================
// forward:  A = softmax_p(Wq X^T) (online), px = A X
// =====================================================================================================================
template <typename T, int E, int NTMAX>
__global__ void __launch_bounds__(NTMAX, 1)
pool_fold_fwd_kernel(const T* __restrict__ X, const float* __restrict__ Wq, float* __restrict__ attn,
                     float* __restrict__ px, const Geom g) {
  pdl_entry();
  constexpr int NV = TG * E;
  MVF_FOLD_SMEM_SETUP();
  float* fin = misc;  // [2][E]: final max, 1/sum

  f2_t wq2[E][4];
  load_vec<E>(Wq + (size_t)g.e0 * g.C + cv * 8, act, wq2, (size_t)g.C);
  const int ridx = Red<NV, 16>::index(lane);
  const int le = lane >> 3, lp = lane & 7;   // the (entity, token-in-group) pair this lane finishes the softmax for

  int n = 0, slot = 0;
  uint32_t phase = 0;
  for (int fi = 0; fi < nF; ++fi) {
    const int64_t f = blockIdx.x + (int64_t)fi * gridDim.x;
    f2_t acc2[E][4];
#pragma unroll
    for (int e = 0; e < E; ++e) acc2[e][0] = acc2[e][1] = acc2[e][2] = acc2[e][3] = 0ull;
    float m_run = -INFINITY, l_run = 0.f;

    for (int gi = 0; gi < nG; ++gi, ++n) {
      const int ntok = min(TG, g.P - gi * TG);
      mbar_wait(&full[slot], phase);
      const uint8_t* st = ring + (size_t)slot * TG * g.row_bytes;
      f2_t x2[TG][4];
      load_group<T>(st, g.row_bytes, cv, act, ntok, x2);
      float part[NV];
      dot_group<E>(wq2, x2, part);   // partial scores of this thread's 8 channels
      Red<NV, 16>::run(part, lane);
      if (ridx >= 0) partial[((n & 1) * NW + warp) * 32 + ridx] = part[0];
      __syncthreads();   // every thread holds its slice of the stage in registers: the slot can be refilled
      if (tid == 0 && n + g.stages < total) issue(n + g.stages);
      if (++slot == g.stages) { slot = 0; phase ^= 1; }

      // softmax bookkeeping, redundantly per warp (lane = (entity, token))
      float s = 0.f;
      if (lane < NV)
        for (int w = 0; w < NW; ++w) s += partial[((n & 1) * NW + w) * 32 + lane];
      const bool valid = lane < NV && lp < ntok;
      if (warp == 0 && valid) table[le * g.P + gi * TG + lp] = s;
      float mx = valid ? s : -INFINITY;
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      const float m_new = fmaxf(m_run, mx);
      const float w_ = valid ? expf(s - m_new) : 0.f;
      const float fac = (lane < NV) ? expf(m_run - m_new) : 0.f;   // first group: exp(-inf) = 0
      float ls = w_;
      ls += __shfl_xor_sync(0xffffffffu, ls, 1);
      ls += __shfl_xor_sync(0xffffffffu, ls, 2);
      ls += __shfl_xor_sync(0xffffffffu, ls, 4);
      l_run = l_run * fac + ls;
      m_run = m_new;
      float* wb = wbuf + warp * 40;
      wb[lane] = w_;
      if (lane < NV && lp == 0) wb[32 + le] = fac;
      __syncwarp();
      // px accumulators: rescale by exp(m_old - m_new), then add this group's weighted tokens
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float fr = wb[32 + e];
        const f2_t f2 = pk2(fr, fr);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc2[e][k] = fmul2(acc2[e][k], f2);
      }
      axpy_group<E>(wb, x2, acc2);
    }

    // ---- end of frame: normalise, write px and the attention map ----
    if (warp == 0 && lane < NV && lp == 0) { fin[le] = m_run; fin[E + le] = 1.f / l_run; }
    __syncthreads();
    if (act) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float inv = fin[E + e];
        float* dst = px + ((size_t)f * g.Etot + g.e0 + e) * g.C + cv * 8;
        const float2 a0 = up2(acc2[e][0]), a1 = up2(acc2[e][1]), a2 = up2(acc2[e][2]), a3 = up2(acc2[e][3]);
        *reinterpret_cast<float4*>(dst) = make_float4(a0.x * inv, a0.y * inv, a1.x * inv, a1.y * inv);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(a2.x * inv, a2.y * inv, a3.x * inv, a3.y * inv);
      }
    }
    for (int i = tid; i < E * g.P; i += blockDim.x) {
      const int e = i / g.P, p = i - e * g.P;
      attn[((size_t)f * g.Etot + g.e0 + e) * g.P + p] = expf(table[i] - fin[e]) * fin[E + e];
    }
    // (the next frame's first write to table / fin happens after that frame's first __syncthreads)
  }
}

// =====================================================================================================================
// backward:  dS = A * (G X^T - delta),  dWq += dS X   (accumulated over every frame of the CTA, one atomic flush)
// =====================================================================================================================
template <typename T, int E, int NTMAX>
__global__ void __launch_bounds__(NTMAX, 1)
pool_fold_bwd_kernel(const T* __restrict__ X, const float* __restrict__ G, const float* __restrict__ px,
                     const float* __restrict__ attn, float* __restrict__ dWq, const Geom g) {
  pdl_entry();
  constexpr int NV = TG * E;
  MVF_FOLD_SMEM_SETUP();
  float* red = misc;  // [NW][E] partial delta

  const int ridx = Red<NV, 16>::index(lane);
  const int le = lane >> 3, lp = lane & 7;
  f2_t acc2[E][4];
#pragma unroll
  for (int e = 0; e < E; ++e) acc2[e][0] = acc2[e][1] = acc2[e][2] = acc2[e][3] = 0ull;

  int n = 0, slot = 0;
  uint32_t phase = 0;
  for (int fi = 0; fi < nF; ++fi) {
    const int64_t f = blockIdx.x + (int64_t)fi * gridDim.x;
    __syncthreads();   // the previous frame's readers of table / red are done
    const size_t off = ((size_t)f * g.Etot + g.e0) * g.C + cv * 8;
    f2_t gv2[E][4];
    load_vec<E>(G + off, act, gv2, (size_t)g.C);
    {
      f2_t pv2[E][4];
      load_vec<E>(px + off, act, pv2, (size_t)g.C);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        f2_t d2 = fmul2(gv2[e][0], pv2[e][0]);
#pragma unroll
        for (int k = 1; k < 4; ++k) d2 = ffma2(gv2[e][k], pv2[e][k], d2);
        const float2 d = up2(d2);
        const float s = warp_sum(d.x + d.y);
        if (lane == 0) red[warp * E + e] = s;
      }
    }
    for (int i = tid; i < E * g.P; i += blockDim.x) {
      const int e = i / g.P, p = i - e * g.P;
      table[i] = attn[((size_t)f * g.Etot + g.e0 + e) * g.P + p];
    }
    __syncthreads();
    float delta = 0.f;
    if (lane < NV)
      for (int w = 0; w < NW; ++w) delta += red[w * E + le];

    for (int gi = 0; gi < nG; ++gi, ++n) {
      const int ntok = min(TG, g.P - gi * TG);
      mbar_wait(&full[slot], phase);
      const uint8_t* st = ring + (size_t)slot * TG * g.row_bytes;
      f2_t x2[TG][4];
      load_group<T>(st, g.row_bytes, cv, act, ntok, x2);
      float part[NV];
      dot_group<E>(gv2, x2, part);
      Red<NV, 16>::run(part, lane);
      if (ridx >= 0) partial[((n & 1) * NW + warp) * 32 + ridx] = part[0];
      __syncthreads();
      if (tid == 0 && n + g.stages < total) issue(n + g.stages);
      if (++slot == g.stages) { slot = 0; phase ^= 1; }

      float dA = 0.f;
      if (lane < NV)
        for (int w = 0; w < NW; ++w) dA += partial[((n & 1) * NW + w) * 32 + lane];
      const bool valid = lane < NV && lp < ntok;
      const float dS = valid ? table[le * g.P + gi * TG + lp] * (dA - delta) : 0.f;
      float* wb = wbuf + warp * 40;
      wb[lane] = dS;
      __syncwarp();
      axpy_group<E>(wb, x2, acc2);
    }
  }
  if (act && nF > 0) {
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = up2(acc2[e][k]);
        atomicAdd(dWq + (size_t)(g.e0 + e) * g.C + cv * 8 + 2 * k, a.x);
        atomicAdd(dWq + (size_t)(g.e0 + e) * g.C + cv * 8 + 2 * k + 1, a.y);
      }
  }
}

// =====================================================================================================================
// small helpers around the streaming passes
// =====================================================================================================================
// Wq[e,c] = scale * sum_j (Q_s[e,j] + Q_b[j]) Wk[j,c];  block = 32 channels x 8 slices of j, deterministic reduction.
// Accumulated in fp64 and rounded once: Wq is shared by every token of every frame, so its rounding error acts like a
// coherent perturbation of W_k (measured: fp32 accumulation costs 1.5e-5 on the gradients at C_in = 1152, fp64 none).
constexpr int PREP_SLICES = 32;   // j-slices per CTA: 32 channels x 32 slices = 1024 threads
__global__ void __launch_bounds__(1024)
fold_prep_kernel(const float* __restrict__ q_s, const float* __restrict__ q_b, const float* __restrict__ Wk, int E, int SPC,
                 int C, float scale, float* __restrict__ Wq) {
  pdl_entry();
  extern __shared__ double smd[];
  float* Qf = reinterpret_cast<float*>(smd + (size_t)PREP_SLICES * E * 32);   // [E][SPC] (fp32 sum, like mvformer.py:383)
  double* red = smd;                                                          // [PREP_SLICES][E][32]
  const int tid = threadIdx.x, cl = tid & 31, js = tid >> 5;
  const int c = blockIdx.x * 32 + cl;
  for (int i = tid; i < E * SPC; i += blockDim.x) Qf[i] = q_s[i] + q_b[i % SPC];
  __syncthreads();
  double acc[MVF_MAX_ENTITIES];
#pragma unroll
  for (int e = 0; e < MVF_MAX_ENTITIES; ++e) acc[e] = 0.0;
  if (c < C) {
#pragma unroll 4
    for (int j = js; j < SPC; j += PREP_SLICES) {
      const double wk = (double)Wk[(size_t)j * C + c];
#pragma unroll
      for (int e = 0; e < MVF_MAX_ENTITIES; ++e)
        if (e < E) acc[e] = fma((double)Qf[e * SPC + j], wk, acc[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < MVF_MAX_ENTITIES; ++e)
    if (e < E) red[(js * E + e) * 32 + cl] = acc[e];
  __syncthreads();
  for (int i = tid; i < E * 32; i += blockDim.x) {
    const int e = i >> 5, l = i & 31;
    double s = 0.0;
#pragma unroll 8
    for (int k = 0; k < PREP_SLICES; ++k) s += red[(k * E + e) * 32 + l];   // fixed order: deterministic
    if (blockIdx.x * 32 + l < C) Wq[(size_t)e * C + blockIdx.x * 32 + l] = (float)(s * (double)scale);
  }
}

// dWk[j,:] += scale * sum_e Qf[e,j] dWq[e,:];  dQ_s[e,j] += scale * dWq[e,:] . Wk[j,:];  dQ_b[j] += sum_e of that
__global__ void __launch_bounds__(256)
fold_finish_kernel(const float* __restrict__ dWq, const float* __restrict__ q_s, const float* __restrict__ q_b,
                   const float* __restrict__ Wk, int E, int SPC, int C, float scale, float* __restrict__ dWk, int64_t ld_dwk,
                   float* __restrict__ dQs, float* __restrict__ dQb) {
  pdl_entry();
  __shared__ float red[8][MVF_MAX_ENTITIES];
  const int j = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float qf[MVF_MAX_ENTITIES], dot[MVF_MAX_ENTITIES];
#pragma unroll
  for (int e = 0; e < MVF_MAX_ENTITIES; ++e) {
    qf[e] = e < E ? (q_s[e * SPC + j] + q_b[j]) * scale : 0.f;
    dot[e] = 0.f;
  }
  if ((C & 3) == 0 && (ld_dwk & 3) == 0 && ((((uintptr_t)Wk) | ((uintptr_t)dWq) | ((uintptr_t)dWk)) & 15) == 0) {
    for (int c = 4 * tid; c < C; c += 4 * blockDim.x) {       // 16-byte path
      const float4 wk = *reinterpret_cast<const float4*>(Wk + (size_t)j * C + c);
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int e = 0; e < MVF_MAX_ENTITIES; ++e) {
        if (e < E) {
          const float4 d = *reinterpret_cast<const float4*>(dWq + (size_t)e * C + c);
          s.x = fmaf(qf[e], d.x, s.x); s.y = fmaf(qf[e], d.y, s.y); s.z = fmaf(qf[e], d.z, s.z); s.w = fmaf(qf[e], d.w, s.w);
          dot[e] = fmaf(d.x, wk.x, fmaf(d.y, wk.y, fmaf(d.z, wk.z, fmaf(d.w, wk.w, dot[e]))));
        }
      }
      float4* o = reinterpret_cast<float4*>(dWk + (size_t)j * ld_dwk + c);
      float4 prev = *o;
      prev.x += s.x; prev.y += s.y; prev.z += s.z; prev.w += s.w;
      *o = prev;
    }
  } else {
    for (int c = tid; c < C; c += blockDim.x) {
      const float wk = Wk[(size_t)j * C + c];
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < MVF_MAX_ENTITIES; ++e) {
        if (e < E) {
          const float d = dWq[(size_t)e * C + c];
          s = fmaf(qf[e], d, s);
          dot[e] = fmaf(d, wk, dot[e]);
        }
      }
      dWk[(size_t)j * ld_dwk + c] += s;
    }
  }
#pragma unroll
  for (int e = 0; e < MVF_MAX_ENTITIES; ++e) {
    if (e < E) {
      const float s = warp_sum(dot[e]);
      if (lane == 0) red[warp][e] = s;
    }
  }
  __syncthreads();
  if (tid < E) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][tid];
    red[0][tid] = s * scale;
  }
  __syncthreads();
  if (tid < E) dQs[tid * SPC + j] += red[0][tid];
  if (tid == 0) {
    float s = 0.f;
    for (int e = 0; e < E; ++e) s += red[0][e];
    dQb[j] += s;
  }
}

// h0[row, :] = [drop(ent[row, :SPC]) | drop(one-hot entity id) | 0 padding]   (mvformer.py:144-151)
__global__ void ent_finish_fwd_kernel(const float* __restrict__ ent, float* __restrict__ h0, int64_t ld, int64_t R, int SPC,
                                      int E, int one_hot, float p, float inv_keep, DropSeed seed) {
  pdl_entry();
  const int W = SPC + (one_hot ? E : 0);
  const int64_t total = R * ld;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / ld;
    const int c = (int)(i - row * ld);
    float v = 0.f;
    if (c < SPC) v = ent[row * SPC + c];
    else if (one_hot && c - SPC == (int)(row % E)) v = 1.f;
    if (v != 0.f && p > 0.f && c < W) v *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + c), p, inv_keep);
    h0[i] = v;
  }
}
// the same, four columns per thread (SPC % 4 == 0, ld % 4 == 0, < 2^31 elements)
__global__ void __launch_bounds__(256)
ent_finish_fwd_v4_kernel(const float* __restrict__ ent, float* __restrict__ h0, int ld4, int R, int SPC, int E, int one_hot,
                         float p, float inv_keep, DropSeed seed) {
  pdl_entry();
  const int W = SPC + (one_hot ? E : 0);
  const int total4 = R * ld4;
  const uint64_t sd = seed.base + (seed.dev ? __ldg(seed.dev) : 0ull);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += gridDim.x * blockDim.x) {
    const int row = q / ld4, c = (q - row * ld4) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c + 3 < SPC) {
      const float4 e4 = *reinterpret_cast<const float4*>(ent + (int64_t)row * SPC + c);
      v[0] = e4.x; v[1] = e4.y; v[2] = e4.z; v[3] = e4.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c + k < SPC) v[k] = ent[(int64_t)row * SPC + c + k];
        else if (one_hot && c + k - SPC == row % E) v[k] = 1.f;
      }
    }
    if (p > 0.f) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (v[k] != 0.f && c + k < W) v[k] *= drop_scale(sd, SITE_FC0, (uint64_t)((int64_t)row * W + c + k), p, inv_keep);
    }
    *reinterpret_cast<float4*>(h0 + (int64_t)q * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// dEnt[row, c] = drop'(d_h0[row, c]) for c < SPC
__global__ void ent_finish_bwd_kernel(const float* __restrict__ d_h0, int64_t ld, float* __restrict__ dEnt, int64_t R, int SPC,
                                      int W, float p, float inv_keep, DropSeed seed) {
  pdl_entry();
  const int64_t total = R * SPC;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / SPC;
    const int c = (int)(i - row * SPC);
    float v = d_h0[row * ld + c];
    if (p > 0.f) v *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + c), p, inv_keep);
    dEnt[i] = v;
  }
}

// same, one warp per row, plus delta[row] = <dEnt[row, :], ent[row, :] - bv>.  With ent = px Wv^T + bv and G = dEnt Wv this
// equals <G_row, px_row>, the softmax-Jacobian constant of the folded pooling backward, at SPC instead of C_in MACs per row
__global__ void __launch_bounds__(256)
ent_finish_bwd_delta_kernel(const float* __restrict__ d_h0, int64_t ld, float* __restrict__ dEnt, int64_t R, int SPC, int W,
                            float p, float inv_keep, DropSeed seed, const float* __restrict__ ent,
                            const float* __restrict__ bv, float* __restrict__ delta) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  float acc = 0.f;
  if ((SPC & 3) == 0 && (ld & 3) == 0 && ((((uintptr_t)d_h0) | ((uintptr_t)dEnt) | ((uintptr_t)ent) | ((uintptr_t)bv)) & 15) == 0) {
    const uint64_t sd = seed.base + (seed.dev ? __ldg(seed.dev) : 0ull);
    for (int c = 4 * lane; c < SPC; c += 128) {              // 16-byte path
      float4 v = *reinterpret_cast<const float4*>(d_h0 + row * ld + c);
      if (p > 0.f) {
        const uint64_t i0 = (uint64_t)(row * W + c);
        v.x *= drop_scale(sd, SITE_FC0, i0, p, inv_keep);
        v.y *= drop_scale(sd, SITE_FC0, i0 + 1, p, inv_keep);
        v.z *= drop_scale(sd, SITE_FC0, i0 + 2, p, inv_keep);
        v.w *= drop_scale(sd, SITE_FC0, i0 + 3, p, inv_keep);
      }
      *reinterpret_cast<float4*>(dEnt + row * SPC + c) = v;
      const float4 en = *reinterpret_cast<const float4*>(ent + row * SPC + c), b4 = *reinterpret_cast<const float4*>(bv + c);
      acc = fmaf(v.x, en.x - b4.x, fmaf(v.y, en.y - b4.y, fmaf(v.z, en.z - b4.z, fmaf(v.w, en.w - b4.w, acc))));
    }
  } else {
    for (int c = lane; c < SPC; c += 32) {
      float v = d_h0[row * ld + c];
      if (p > 0.f) v *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + c), p, inv_keep);
      dEnt[row * SPC + c] = v;
      acc = fmaf(v, ent[row * SPC + c] - bv[c], acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) delta[row] = acc;
}

// ---- launch plumbing ---------------------------------------------------------------------------------------------------
static int g_sms = -1;
static int num_sms() {
  if (g_sms < 0) {
    int dev = 0;
    cudaDeviceProp prop;
    g_sms = (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) ? prop.multiProcessorCount : 1;
  }
  return g_sms;
}

struct LaunchCfg {
  int nt, stages, grid;
  size_t smem;
};

// Planned once per (kernel instantiation, shape): cudaFuncSetAttribute / the occupancy query are host-side calls that
// must not sit on the per-step launch path.
struct PlanKey {
  int F, P, C;
  bool valid = false;
  LaunchCfg cfg;
};
template <typename KernelT>
static int plan_launch(KernelT kernel, PlanKey& cache, int F, int P, int C, int E, int elem, LaunchCfg& L) {
  if (cache.valid && cache.F == F && cache.P == P && cache.C == C) {
    L = cache.cfg;
    return MVF_OK;
  }
  L.nt = (int)round_up(C / 8, 32);
  const int row_bytes = C * elem, nw = L.nt / 32;
  int stages = 4;
  const char* es = getenv("MVF_FOLD_STAGES");
  if (es && atoi(es) >= 2 && atoi(es) <= MAX_STAGES) stages = atoi(es);
  while (stages > 2 && carve(stages, row_bytes, nw, E, P).total > 200 * 1024) --stages;
  Carve cv = carve(stages, row_bytes, nw, E, P);
  MVF_REQUIRE(cv.total <= 227 * 1024, MVF_ERR_UNSUPPORTED, "pool_fold: %d channels x %d B need %zu B of shared memory", C,
              elem, cv.total);
  L.stages = stages;
  L.smem = cv.total;
  MVF_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  int occ = 0;
  MVF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, L.nt, L.smem));
  if (occ < 1) occ = 1;
  const char* eo = getenv("MVF_FOLD_CTAS_PER_SM");
  if (eo && atoi(eo) > 0 && atoi(eo) < occ) occ = atoi(eo);
  const int64_t slots = (int64_t)occ * num_sms();
  L.grid = (int)(F < slots ? F : slots);
  cache.F = F; cache.P = P; cache.C = C; cache.cfg = L; cache.valid = true;
  return MVF_OK;
}

template <typename T, int E, int NTMAX>
static int fwd_launch(const Geom& g, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  LaunchCfg L;
  static thread_local PlanKey cache;
  MVF_TRY(plan_launch(pool_fold_fwd_kernel<T, E, NTMAX>, cache, g.F, g.P, g.C, E, (int)sizeof(T), L));
  Geom gg = g;
  gg.stages = L.stages;
  launch_k(pool_fold_fwd_kernel<T, E, NTMAX>, L.grid, L.nt, L.smem, st, (const T*)X, Wq, attn, px, gg);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
template <typename T, int E, int NTMAX>
static int bwd_launch(const Geom& g, const void* X, const float* G, const float* px, const float* attn, float* dWq,
                      cudaStream_t st) {
  LaunchCfg L;
  static thread_local PlanKey cache;
  MVF_TRY(plan_launch(pool_fold_bwd_kernel<T, E, NTMAX>, cache, g.F, g.P, g.C, E, (int)sizeof(T), L));
  Geom gg = g;
  gg.stages = L.stages;
  launch_k(pool_fold_bwd_kernel<T, E, NTMAX>, L.grid, L.nt, L.smem, st, (const T*)X, G, px, attn, dWq, gg);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

template <typename T, int E>
static int fwd_nt(const Geom& g, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  const int nt = g.C / 8;
  if (nt <= 128) return fwd_launch<T, E, 128>(g, X, Wq, attn, px, st);
  if (nt <= 320) return fwd_launch<T, E, 320>(g, X, Wq, attn, px, st);
  return fwd_launch<T, E, 640>(g, X, Wq, attn, px, st);
}
template <typename T, int E>
static int bwd_nt(const Geom& g, const void* X, const float* G, const float* px, const float* attn, float* dWq, cudaStream_t st) {
  const int nt = g.C / 8;
  if (nt <= 128) return bwd_launch<T, E, 128>(g, X, G, px, attn, dWq, st);
  if (nt <= 320) return bwd_launch<T, E, 320>(g, X, G, px, attn, dWq, st);
  return bwd_launch<T, E, 640>(g, X, G, px, attn, dWq, st);
}
template <typename T>
static int fwd_e(int e, const Geom& g, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  switch (e) {
    case 1: return fwd_nt<T, 1>(g, X, Wq, attn, px, st);
    case 2: return fwd_nt<T, 2>(g, X, Wq, attn, px, st);
    case 3: return fwd_nt<T, 3>(g, X, Wq, attn, px, st);
    default: return fwd_nt<T, 4>(g, X, Wq, attn, px, st);
  }
}
template <typename T>
static int bwd_e(int e, const Geom& g, const void* X, const float* G, const float* px, const float* attn, float* dWq,
                 cudaStream_t st) {
  switch (e) {
    case 1: return bwd_nt<T, 1>(g, X, G, px, attn, dWq, st);
    case 2: return bwd_nt<T, 2>(g, X, G, px, attn, dWq, st);
    case 3: return bwd_nt<T, 3>(g, X, G, px, attn, dWq, st);
    default: return bwd_nt<T, 4>(g, X, G, px, attn, dWq, st);
  }
}

}  // namespace fold

bool pool_fold_supported(int dtype, int C, int E, int P) {
  (void)dtype;
  return C > 0 && C % 8 == 0 && C / 8 <= 640 && E >= 1 && E <= MVF_MAX_ENTITIES && P >= 1;
}

// bf16 tokens with C_in % 16 == 0 run on the warp-specialised tensor-core kernels of pool_fold_ws.cu; MVF_FOLD_WS=0 forces the
// CUDA-core kernels of this file (A/B measurements, and the path fp32 tokens always take)
static bool use_ws(int dtype, int C, int P) {
  const char* e = getenv("MVF_FOLD_WS");
  if (e && atoi(e) == 0) return false;
  return pool_fold_ws_supported(dtype, C, P);
}

static int fold_check(int dtype, int F, int P, int E, int C, const void* X) {
  MVF_REQUIRE(pool_fold_supported(dtype, C, E, P), MVF_ERR_UNSUPPORTED,
              "pool_fold: needs C_in (%d) a multiple of 8 and at most 5120, 1 <= E (%d) <= %d", C, E, MVF_MAX_ENTITIES);
  MVF_REQUIRE(F >= 0 && X != nullptr && (((uintptr_t)X) & 15) == 0, MVF_ERR_ALIGN, "pool_fold: tokens must be 16-byte aligned");
  return MVF_OK;
}

int pool_fold_fwd(int dtype, int F, int P, int E, int C, const void* X, const float* Wq, float* attn, float* px,
                  cudaStream_t st) {
  MVF_TRY(fold_check(dtype, F, P, E, C, X));
  if (F == 0) return MVF_OK;
  if (use_ws(dtype, C, P)) return pool_fold_ws_fwd(F, P, E, C, X, Wq, attn, px, st);
  for (int e0 = 0; e0 < E; e0 += 4) {   // entity passes of <= 4 (X is re-streamed per pass; E = 3 in every penn config)
    fold::Geom g{F, P, C, E, e0, 0, C * (dtype == MVF_BF16 ? 2 : 4)};
    const int ne = E - e0 < 4 ? E - e0 : 4;
    if (dtype == MVF_BF16) MVF_TRY(fold::fwd_e<bf16>(ne, g, X, Wq, attn, px, st));
    else MVF_TRY(fold::fwd_e<float>(ne, g, X, Wq, attn, px, st));
  }
  return MVF_OK;
}

int pool_fold_bwd(int dtype, int F, int P, int E, int C, const void* X, const float* G, const float* px, const float* attn,
                  float* dWq, cudaStream_t st, const float* delta) {
  MVF_TRY(fold_check(dtype, F, P, E, C, X));
  if (F == 0) return MVF_OK;
  if (use_ws(dtype, C, P)) return pool_fold_ws_bwd(F, P, E, C, X, G, px, attn, delta, dWq, st);
  for (int e0 = 0; e0 < E; e0 += 4) {
    fold::Geom g{F, P, C, E, e0, 0, C * (dtype == MVF_BF16 ? 2 : 4)};
    const int ne = E - e0 < 4 ? E - e0 : 4;
    if (dtype == MVF_BF16) MVF_TRY(fold::bwd_e<bf16>(ne, g, X, G, px, attn, dWq, st));
    else MVF_TRY(fold::bwd_e<float>(ne, g, X, G, px, attn, dWq, st));
  }
  return MVF_OK;
}

int fold_prep(const float* q_s, const float* q_b, const float* Wk, int E, int SPC, int C, float* Wq, cudaStream_t st) {
  const size_t smem = (size_t)E * SPC * sizeof(float) + (size_t)fold::PREP_SLICES * E * 32 * sizeof(double);
  MVF_REQUIRE(smem <= 200 * 1024, MVF_ERR_UNSUPPORTED, "fold_prep: E*SPC too large");
  if (smem > 48 * 1024)
    MVF_CHECK_CUDA(cudaFuncSetAttribute(fold::fold_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_k(fold::fold_prep_kernel, cdiv(C, 32), 1024, smem, st, q_s, q_b, Wk, E, SPC, C, (float)(1.0 / sqrt((double)SPC)), Wq);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

int fold_finish(const float* dWq, const float* q_s, const float* q_b, const float* Wk, int E, int SPC, int C, float* dWk,
                int64_t ld_dwk, float* dQs, float* dQb, cudaStream_t st) {
  launch_k(fold::fold_finish_kernel, SPC, 256, 0, st, dWq, q_s, q_b, Wk, E, SPC, C, 1.f / sqrtf((float)SPC), dWk, ld_dwk, dQs, dQb);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

int ent_finish_fwd(const float* ent, float* h0, int64_t ld, int64_t R, int SPC, int E, int one_hot, float p, DropSeed seed,
                   cudaStream_t st) {
  if (R <= 0) return MVF_OK;
  const int64_t total = R * ld;
  if ((SPC & 3) == 0 && (ld & 3) == 0 && total < (1ll << 31) && ((((uintptr_t)ent) | ((uintptr_t)h0)) & 15) == 0) {
    const int64_t work = total / 4;
    launch_k(fold::ent_finish_fwd_v4_kernel, (int)((work + 255) / 256 < 2368 ? (work + 255) / 256 : 2368), 256, 0, st, ent, h0,
             (int)(ld / 4), (int)R, SPC, E, one_hot, p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  }
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  launch_k(fold::ent_finish_fwd_kernel, blocks, 256, 0, st, ent, h0, ld, R, SPC, E, one_hot, p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

int ent_finish_bwd(const float* d_h0, int64_t ld, float* dEnt, int64_t R, int SPC, int W, float p, DropSeed seed,
                   cudaStream_t st, const float* ent, const float* bv, float* delta) {
  if (R <= 0) return MVF_OK;
  if (delta != nullptr && ent != nullptr && bv != nullptr) {
    launch_k(fold::ent_finish_bwd_delta_kernel, (int)((R + 7) / 8), 256, 0, st, d_h0, ld, dEnt, R, SPC, W, p,
                                                                         p > 0.f ? 1.f / (1.f - p) : 1.f, seed, ent, bv, delta);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  }
  const int64_t total = R * SPC;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  launch_k(fold::ent_finish_bwd_kernel, blocks, 256, 0, st, d_h0, ld, dEnt, R, SPC, W, p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace mvf

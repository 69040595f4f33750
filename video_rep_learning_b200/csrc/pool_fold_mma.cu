// Folded entity pooling for bf16 tokens on the warp-level tensor cores (mma.sync m16n8k16, fp32 accumulate).
//
// Same mathematics and the same single streaming pass over the tokens as pool_fold.cu (see the header there); what
// changes is who does the two skinny contractions of a 16-token group.  With CUDA-core FMAs the kernel is bound by
// instruction issue (measured 40 % of HBM peak: ~750 instructions per thread per 8 tokens, 9 warps per SM); here
//
//   scores   S[16 tok, 8 ent]   = X[16 tok, C]      * Wq^T[C, 8 ent]       A = X tile (ldmatrix),       B = Wq  as bf16 hi+lo
//   pooling  px^T[C, 8 ent]    += X^T[C, 16 tok]    * w^T[16 tok, 8 ent]   A = X tile (ldmatrix.trans), B = w   as bf16 hi+lo
//
// each warp owns TPW 16-channel tiles: 2*TPW ldmatrix + 4*TPW mma per 16 tokens instead of ~1500 FMA-path instructions.
// The tokens are bf16 already, so A is exact; the fp32 B operands (Wq or G, and the softmax weights) are split into
// bf16 hi + lo and multiplied in two MMAs (16 mantissa bits).  This is HBM-bound integer-free streaming work, not a
// GEMM reshaped for the tensor cores: N = 8 is the entity count padded to the MMA shape, and the kernel's roofline
// stays the HBM read of X.
//
// CTA = NW compute warps + 1 producer warp.  The producer streams 16-token groups into a 2-slot shared-memory ring
// with one bulk async copy per token row (rows are padded by 16 B in shared memory so that ldmatrix is conflict-free)
// and full/empty mbarriers; compute warps meet once per group at a named barrier (cross-warp reduction of S).
#include <math.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace mvf {
namespace foldm {

constexpr int TG = 16;       // tokens per group
constexpr int SLOTS = 2;
constexpr int EN = 8;        // entities per pass (MMA N)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
      printf("mvf pool_fold_mma: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void compute_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&a)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&a)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x, y) -> bf16 hi pair and bf16 lo pair (x = hi + lo up to 2^-17)
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// register budget: 12 (8) tiles per warp hold 48 (32) B-fragment + 48 (32) accumulator registers -> at most 12 compute warps
__host__ __device__ constexpr int max_compute_warps(int tpw) { return tpw >= 8 ? 12 : 16; }
__host__ __device__ constexpr int max_threads(int tpw) { return (max_compute_warps(tpw) + 1) * 32; }

struct Geom {
  int F, P, C;
  int Etot, e0, ne;   // entities in the model, first entity of this pass, entities in this pass (<= 8)
  int NW;             // compute warps
  int pitch;          // shared-memory row pitch in bytes: C*2 + 16
};

struct Carve {
  size_t ring, bars, partial, wbuf, table, misc, total;
};
static Carve carve(int C, int NW, int P) {
  Carve c;
  c.ring = (size_t)SLOTS * TG * (C * 2 + 16);
  c.bars = 64;
  c.partial = (size_t)2 * NW * EN * TG * 4;
  c.wbuf = (size_t)NW * (EN * TG + 16) * 4;
  c.table = ((size_t)EN * P * 4 + 15) / 16 * 16;
  c.misc = (size_t)NW * 16 * 4;
  c.total = c.ring + c.bars + c.partial + c.wbuf + c.table + c.misc;
  return c;
}

#define MVF_FOLDM_SETUP()                                                                              \
  extern __shared__ __align__(128) uint8_t smraw[];                                                    \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = g.NW;                             \
  uint8_t* ring = smraw;                                                                               \
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)SLOTS * TG * g.pitch);                   \
  uint64_t* empty = full + SLOTS;                                                                      \
  float* partial = reinterpret_cast<float*>(full + 8);                                                 \
  float* wbuf = partial + 2 * NW * EN * TG;                                                            \
  float* table = wbuf + NW * (EN * TG + 16);                                                           \
  float* misc = table + (((size_t)EN * g.P + 3) / 4) * 4;                                              \
  const int nG = (g.P + TG - 1) / TG;                                                                  \
  const int nF = blockIdx.x < g.F ? (g.F - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;                \
  const int total = nF * nG;                                                                           \
  if (tid == 0) {                                                                                      \
    for (int s = 0; s < SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }              \
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                                 \
  }                                                                                                    \
  __syncthreads();                                                                                     \
  if (warp == NW) {                                                                                    \
    /* ===== producer warp: one bulk copy per token row; rows past the end of the frame re-read valid rows ===== */ \
    if (lane == 0) {                                                                                   \
      const uint32_t row_bytes = (uint32_t)g.C * 2u;                                                   \
      for (int n = 0; n < total; ++n) {                                                                \
        const int s = n % SLOTS;                                                                       \
        if (n >= SLOTS) mbar_wait(&empty[s], ((n / SLOTS) - 1) & 1);                                   \
        const int fi = n / nG, gi = n - fi * nG;                                                       \
        const int64_t f = blockIdx.x + (int64_t)fi * gridDim.x;                                        \
        mbar_expect_tx(&full[s], TG * row_bytes);                                                      \
        const uint8_t* src = reinterpret_cast<const uint8_t*>(X) + f * g.P * (int64_t)row_bytes;       \
        uint8_t* dst = ring + (size_t)s * TG * g.pitch;                                                \
        for (int r = 0; r < TG; ++r) {                                                                 \
          int tok = gi * TG + r;                                                                       \
          if (tok >= g.P) tok %= g.P;                                                                  \
          bulk_g2s(dst + (size_t)r * g.pitch, src + (int64_t)tok * row_bytes, row_bytes, &full[s]);    \
        }                                                                                              \
      }                                                                                                \
    }                                                                                                  \
    return;                                                                                            \
  }                                                                                                    \
  const int q = lane & 3, r8 = lane >> 2;                /* MMA fragment coordinates */                \
  const int ltok = lane & 15, lh = lane >> 4;            /* softmax bookkeeping: token, entity parity */ \
  const int nthr = NW * 32;                                                                            \
  const uint32_t ring_u32 = smem_u32(ring);                                                            \
  /* per-lane ldmatrix byte offsets inside a slot (tile 0 of this warp) */                             \
  const uint32_t a_off = (uint32_t)((lane & 15) * g.pitch + (warp * TPW * 16 + (lane >> 4) * 8) * 2);  \
  const uint32_t t_off = (uint32_t)(((lane & 7) + 8 * (lane >> 4)) * g.pitch + (warp * TPW * 16 + ((lane >> 3) & 1) * 8) * 2);

// bf16 hi/lo B fragments of a [8 ent, C] fp32 matrix (rows e >= ne are zero) for this warp's tiles
template <int TPW>
__device__ __forceinline__ void load_bfrag(const float* __restrict__ M, int64_t stride, int ne, int ch_base, int q, int r8,
                                           uint32_t (&bh)[TPW][2], uint32_t (&bl)[TPW][2], float mul) {
#pragma unroll
  for (int t = 0; t < TPW; ++t) {
    float2 v0 = make_float2(0.f, 0.f), v1 = v0;
    if (r8 < ne) {
      const float* p = M + (int64_t)r8 * stride + ch_base + t * 16 + 2 * q;
      v0 = *reinterpret_cast<const float2*>(p);
      v1 = *reinterpret_cast<const float2*>(p + 8);
    }
    split2(v0.x * mul, v0.y * mul, bh[t][0], bl[t][0]);
    split2(v1.x * mul, v1.y * mul, bh[t][1], bl[t][1]);
  }
}

// S[16 tok, 8 ent] partial of this warp's channels -> partial[buf][warp][ent][tok]
template <int TPW>
__device__ __forceinline__ void scores_phase(uint32_t slot_addr, uint32_t a_off, const uint32_t (&bh)[TPW][2],
                                             const uint32_t (&bl)[TPW][2], float* pw, int q, int r8) {
  // four independent accumulation chains (lo / hi operand x even / odd tile): a single chain would serialise 2*TPW
  // dependent MMAs per group (measured: "wait" + "short scoreboard" stalls dominate, tensor pipe 22 % busy)
  float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f}, sd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < TPW; ++t) {
    uint32_t a[4];
    ldsm_x4(slot_addr + a_off + t * 32, a);
    if (t & 1) {
      mma16816(sc, a, bl[t][0], bl[t][1]);
      mma16816(sd, a, bh[t][0], bh[t][1]);
    } else {
      mma16816(sa, a, bl[t][0], bl[t][1]);
      mma16816(sb, a, bh[t][0], bh[t][1]);
    }
  }
  float s[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = (sa[i] + sc[i]) + (sb[i] + sd[i]);
  pw[(2 * q) * TG + r8] = s[0];
  pw[(2 * q + 1) * TG + r8] = s[1];
  pw[(2 * q) * TG + r8 + 8] = s[2];
  pw[(2 * q + 1) * TG + r8 + 8] = s[3];
}

// acc[t] (= px^T tile [16 ch, 8 ent]) += X^T tile * w^T, w[ent][tok] read from this warp's wbuf
template <int TPW>
__device__ __forceinline__ void pool_phase(uint32_t slot_addr, uint32_t t_off, const float* wb, int ne, float (&acc)[TPW][4],
                                           int q, int r8) {
  float2 w0 = make_float2(0.f, 0.f), w1 = w0;
  if (r8 < ne) {
    w0 = *reinterpret_cast<const float2*>(wb + r8 * TG + 2 * q);
    w1 = *reinterpret_cast<const float2*>(wb + r8 * TG + 2 * q + 8);
  }
  uint32_t wh[2], wl[2];
  split2(w0.x, w0.y, wh[0], wl[0]);
  split2(w1.x, w1.y, wh[1], wl[1]);
#pragma unroll
  for (int t = 0; t < TPW; ++t) {
    uint32_t a[4];
    ldsm_x4_trans(slot_addr + t_off + t * 32, a);
    mma16816(acc[t], a, wl[0], wl[1]);
    mma16816(acc[t], a, wh[0], wh[1]);
  }
}

// =====================================================================================================================
// forward
// =====================================================================================================================
// NI = ceil(ne / 2): rounds of the softmax bookkeeping (each round serves two entities, one per half-warp)
template <int TPW, int NI>
__global__ void __launch_bounds__(max_threads(TPW), 1)
pool_foldm_fwd_kernel(const bf16* __restrict__ X, const float* __restrict__ Wq, float* __restrict__ attn,
                      float* __restrict__ px, const Geom g) {
  pdl_entry();
  MVF_FOLDM_SETUP();
  float* fin = misc + warp * 16;   // this warp's copy of [max(8) | 1/sum(8)] at the end of a frame

  // the scores live in the log2 domain (log2 e folded into the Wq fragments): exp2f is one MUFU + fix-ups, expf ~4x that
  uint32_t bh[TPW][2], bl[TPW][2];
  load_bfrag<TPW>(Wq + (int64_t)g.e0 * g.C, g.C, g.ne, warp * TPW * 16, q, r8, bh, bl, 1.4426950408889634f);

  int n = 0;
  for (int fi = 0; fi < nF; ++fi) {
    const int64_t f = blockIdx.x + (int64_t)fi * gridDim.x;
    float acc[TPW][4];
#pragma unroll
    for (int t = 0; t < TPW; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
    float m_run[NI], l_run[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) { m_run[i] = -INFINITY; l_run[i] = 0.f; }

    for (int gi = 0; gi < nG; ++gi, ++n) {
      const int slot = n % SLOTS;
      const int ntok = min(TG, g.P - gi * TG);
      mbar_wait(&full[slot], (n / SLOTS) & 1);
      const uint32_t slot_addr = ring_u32 + (uint32_t)(slot * TG * g.pitch);
      scores_phase<TPW>(slot_addr, a_off, bh, bl, partial + ((n & 1) * NW + warp) * (EN * TG), q, r8);
      compute_sync(nthr);

      // softmax bookkeeping, redundantly per warp: lane = (token ltok, entities lh, lh+2, ...).  The NI rounds are written
      // as straight-line code over small arrays so that their shuffle / exp2 chains interleave (ILP), not as a loop.
      float* wb = wbuf + warp * (EN * TG + 16);
      const bool tvalid = ltok < ntok;
      float sv[NI], mx[NI], wg[NI], fc[NI], ls[NI];
      bool ev[NI];
      const float* pbase = partial + (n & 1) * NW * (EN * TG) + ltok;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        ev[i] = lh + 2 * i < g.ne;   // the upper half-warp may hold a padding entity: it still takes part in the shuffles
        sv[i] = 0.f;
      }
      for (int w = 0; w < NW; ++w) {
#pragma unroll
        for (int i = 0; i < NI; ++i)
          if (ev[i]) sv[i] += pbase[w * (EN * TG) + (lh + 2 * i) * TG];
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (warp == 0 && tvalid && ev[i]) table[(lh + 2 * i) * g.P + gi * TG + ltok] = sv[i];
        mx[i] = (tvalid && ev[i]) ? sv[i] : -INFINITY;
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
#pragma unroll
        for (int i = 0; i < NI; ++i) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], o));
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float m_new = fmaxf(m_run[i], mx[i]);
        wg[i] = (tvalid && ev[i]) ? exp2f(sv[i] - m_new) : 0.f;
        fc[i] = ev[i] ? exp2f(m_run[i] - m_new) : 1.f;   // first group: 2^-inf = 0
        ls[i] = wg[i];
        if (ev[i]) m_run[i] = m_new;
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
#pragma unroll
        for (int i = 0; i < NI; ++i) ls[i] += __shfl_xor_sync(0xffffffffu, ls[i], o);
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (ev[i]) {
          l_run[i] = l_run[i] * fc[i] + ls[i];
          wb[(lh + 2 * i) * TG + ltok] = wg[i];
          if (ltok == 0) wb[EN * TG + lh + 2 * i] = fc[i];
        }
      }
      __syncwarp();
      // rescale the accumulators when a running maximum moved (columns = entities 2q, 2q+1)
      const float f0 = (2 * q < g.ne) ? wb[EN * TG + 2 * q] : 1.f;
      const float f1 = (2 * q + 1 < g.ne) ? wb[EN * TG + 2 * q + 1] : 1.f;
      if (__any_sync(0xffffffffu, f0 != 1.f || f1 != 1.f)) {
#pragma unroll
        for (int t = 0; t < TPW; ++t) { acc[t][0] *= f0; acc[t][1] *= f1; acc[t][2] *= f0; acc[t][3] *= f1; }
      }
      pool_phase<TPW>(slot_addr, t_off, wb, g.ne, acc, q, r8);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }

    // ---- end of frame ----
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int e = lh + 2 * i;
      if (e < g.ne && ltok == 0) { fin[e] = m_run[i]; fin[8 + e] = 1.f / l_run[i]; }
    }
    __syncwarp();
    {
      const int e0c = 2 * q, e1c = 2 * q + 1;
      const float i0 = e0c < g.ne ? fin[8 + e0c] : 0.f, i1 = e1c < g.ne ? fin[8 + e1c] : 0.f;
      float* base = px + ((int64_t)f * g.Etot + g.e0) * g.C + warp * TPW * 16;
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        const int ch = t * 16 + r8;
        if (e0c < g.ne) { base[(int64_t)e0c * g.C + ch] = acc[t][0] * i0; base[(int64_t)e0c * g.C + ch + 8] = acc[t][2] * i0; }
        if (e1c < g.ne) { base[(int64_t)e1c * g.C + ch] = acc[t][1] * i1; base[(int64_t)e1c * g.C + ch + 8] = acc[t][3] * i1; }
      }
    }
    compute_sync(nthr);   // warp 0's raw scores of the whole frame are visible
    for (int i = tid; i < g.ne * g.P; i += nthr) {
      const int e = i / g.P, p = i - e * g.P;
      attn[((int64_t)f * g.Etot + g.e0 + e) * g.P + p] = exp2f(table[i] - fin[e]) * fin[8 + e];
    }
    compute_sync(nthr);   // table / fin are free for the next frame
  }
}

// =====================================================================================================================
// backward
// =====================================================================================================================
template <int TPW>
__global__ void __launch_bounds__(max_threads(TPW), 1)
pool_foldm_bwd_kernel(const bf16* __restrict__ X, const float* __restrict__ G, const float* __restrict__ px,
                      const float* __restrict__ attn, float* __restrict__ dWq, const Geom g) {
  pdl_entry();
  MVF_FOLDM_SETUP();
  float* red = misc;   // [NW][8] partial delta (indexing: warp * 16 + e)

  float acc[TPW][4];
#pragma unroll
  for (int t = 0; t < TPW; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

  int n = 0;
  for (int fi = 0; fi < nF; ++fi) {
    const int64_t f = blockIdx.x + (int64_t)fi * gridDim.x;
    const int64_t rowbase = ((int64_t)f * g.Etot + g.e0) * g.C;
    uint32_t bh[TPW][2], bl[TPW][2];
    float dpart = 0.f;
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      float2 v0 = make_float2(0.f, 0.f), v1 = v0;
      if (r8 < g.ne) {
        const int64_t o = rowbase + (int64_t)r8 * g.C + warp * TPW * 16 + t * 16 + 2 * q;
        v0 = *reinterpret_cast<const float2*>(G + o);
        v1 = *reinterpret_cast<const float2*>(G + o + 8);
        const float2 p0 = *reinterpret_cast<const float2*>(px + o), p1 = *reinterpret_cast<const float2*>(px + o + 8);
        dpart += v0.x * p0.x + v0.y * p0.y + v1.x * p1.x + v1.y * p1.y;
      }
      split2(v0.x, v0.y, bh[t][0], bl[t][0]);
      split2(v1.x, v1.y, bh[t][1], bl[t][1]);
    }
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 1);
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 2);
    if (q == 0) red[warp * 16 + r8] = dpart;     // entity r8 (zero for r8 >= ne)
    for (int i = tid; i < g.ne * g.P; i += nthr) {
      const int e = i / g.P, p = i - e * g.P;
      table[i] = attn[((int64_t)f * g.Etot + g.e0 + e) * g.P + p];
    }
    compute_sync(nthr);
    float delta[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      delta[i] = 0.f;
      const int e = lh + 2 * i;
      if (e < g.ne)
        for (int w = 0; w < NW; ++w) delta[i] += red[w * 16 + e];
    }

    for (int gi = 0; gi < nG; ++gi, ++n) {
      const int slot = n % SLOTS;
      const int ntok = min(TG, g.P - gi * TG);
      mbar_wait(&full[slot], (n / SLOTS) & 1);
      const uint32_t slot_addr = ring_u32 + (uint32_t)(slot * TG * g.pitch);
      scores_phase<TPW>(slot_addr, a_off, bh, bl, partial + ((n & 1) * NW + warp) * (EN * TG), q, r8);
      compute_sync(nthr);
      float* wb = wbuf + warp * (EN * TG + 16);
      const bool tvalid = ltok < ntok;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = lh + 2 * i;
        if (e < g.ne) {
          float dA = 0.f;
          for (int w = 0; w < NW; ++w) dA += partial[((n & 1) * NW + w) * (EN * TG) + e * TG + ltok];
          wb[e * TG + ltok] = tvalid ? table[e * g.P + gi * TG + ltok] * (dA - delta[i]) : 0.f;
        }
      }
      __syncwarp();
      pool_phase<TPW>(slot_addr, t_off, wb, g.ne, acc, q, r8);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    compute_sync(nthr);   // table / red are free for the next frame
  }
  if (nF > 0) {
    const int e0c = 2 * q, e1c = 2 * q + 1;
    float* base = dWq + (int64_t)g.e0 * g.C + warp * TPW * 16;
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const int ch = t * 16 + r8;
      if (e0c < g.ne) { atomicAdd(base + (int64_t)e0c * g.C + ch, acc[t][0]); atomicAdd(base + (int64_t)e0c * g.C + ch + 8, acc[t][2]); }
      if (e1c < g.ne) { atomicAdd(base + (int64_t)e1c * g.C + ch, acc[t][1]); atomicAdd(base + (int64_t)e1c * g.C + ch + 8, acc[t][3]); }
    }
  }
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

// tiles per warp / compute warps for C channels (C % 16 == 0): prefer a warp count that is a multiple of 4 (one CTA per
// SM, 4 schedulers), then more warps; at most 16 compute warps and 12 tiles per warp (register budget)
static bool choose_shape(int C, int* tpw_out, int* nw_out) {
  const int CT = C / 16;
  static const int cand[] = {12, 8, 6, 4, 3, 2, 1};
  int best_tpw = 0, best_nw = 0, best_score = -1;
  for (int tpw : cand) {
    if (CT % tpw) continue;
    const int nw = CT / tpw;
    if (nw > max_compute_warps(tpw)) continue;
    const int score = (nw % 4 == 0 ? 100 : 0) + nw;
    if (score > best_score) { best_score = score; best_tpw = tpw; best_nw = nw; }
  }
  if (best_score < 0) return false;
  *tpw_out = best_tpw;
  *nw_out = best_nw;
  return true;
}

struct Plan {
  int F = -1, P = -1, C = -1;
  int grid = 0;
  size_t smem = 0;
};

template <typename KernelT>
static int plan(KernelT kernel, Plan& pl, int F, int P, int C, int NW) {
  if (pl.F == F && pl.P == P && pl.C == C) return MVF_OK;
  Carve cv = carve(C, NW, P);
  MVF_REQUIRE(cv.total <= 227 * 1024, MVF_ERR_UNSUPPORTED, "pool_fold_mma: %d channels need %zu B of shared memory", C, cv.total);
  MVF_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cv.total));
  int occ = 0;
  MVF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, (NW + 1) * 32, cv.total));
  if (occ < 1) occ = 1;
  const int64_t slots = (int64_t)occ * num_sms();
  pl.grid = (int)(F < slots ? F : slots);
  pl.smem = cv.total;
  pl.F = F; pl.P = P; pl.C = C;
  return MVF_OK;
}

template <int TPW, int NI>
static int fwd_launch_ni(const Geom& g, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  static thread_local Plan pl;
  MVF_TRY(plan(pool_foldm_fwd_kernel<TPW, NI>, pl, g.F, g.P, g.C, g.NW));
  launch_k(pool_foldm_fwd_kernel<TPW, NI>, pl.grid, (g.NW + 1) * 32, pl.smem, st, (const bf16*)X, Wq, attn, px, g);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
template <int TPW>
static int fwd_launch(const Geom& g, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  switch ((g.ne + 1) / 2) {
    case 1: return fwd_launch_ni<TPW, 1>(g, X, Wq, attn, px, st);
    case 2: return fwd_launch_ni<TPW, 2>(g, X, Wq, attn, px, st);
    case 3: return fwd_launch_ni<TPW, 3>(g, X, Wq, attn, px, st);
    default: return fwd_launch_ni<TPW, 4>(g, X, Wq, attn, px, st);
  }
}
template <int TPW>
static int bwd_launch(const Geom& g, const void* X, const float* G, const float* px, const float* attn, float* dWq,
                      cudaStream_t st) {
  static thread_local Plan pl;
  MVF_TRY(plan(pool_foldm_bwd_kernel<TPW>, pl, g.F, g.P, g.C, g.NW));
  launch_k(pool_foldm_bwd_kernel<TPW>, pl.grid, (g.NW + 1) * 32, pl.smem, st, (const bf16*)X, G, px, attn, dWq, g);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace foldm

bool pool_fold_mma_supported(int dtype, int C, int P) {
  int tpw, nw;
  if (dtype != MVF_BF16 || C <= 0 || C % 16 != 0 || P < 1) return false;
  if (!foldm::choose_shape(C, &tpw, &nw)) return false;
  return foldm::carve(C, nw, P).total <= 227 * 1024;
}

#define MVF_FOLDM_DISPATCH(FN, ...)                      \
  switch (tpw) {                                         \
    case 1: return foldm::FN<1>(__VA_ARGS__);            \
    case 2: return foldm::FN<2>(__VA_ARGS__);            \
    case 3: return foldm::FN<3>(__VA_ARGS__);            \
    case 4: return foldm::FN<4>(__VA_ARGS__);            \
    case 6: return foldm::FN<6>(__VA_ARGS__);            \
    case 8: return foldm::FN<8>(__VA_ARGS__);            \
    default: return foldm::FN<12>(__VA_ARGS__);          \
  }

static int foldm_fwd_pass(const foldm::Geom& g, int tpw, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  MVF_FOLDM_DISPATCH(fwd_launch, g, X, Wq, attn, px, st)
}
static int foldm_bwd_pass(const foldm::Geom& g, int tpw, const void* X, const float* G, const float* px, const float* attn,
                          float* dWq, cudaStream_t st) {
  MVF_FOLDM_DISPATCH(bwd_launch, g, X, G, px, attn, dWq, st)
}

int pool_fold_mma_fwd(int F, int P, int E, int C, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  int tpw = 0, nw = 0;
  MVF_REQUIRE(foldm::choose_shape(C, &tpw, &nw), MVF_ERR_UNSUPPORTED, "pool_fold_mma: unsupported channel count %d", C);
  for (int e0 = 0; e0 < E; e0 += foldm::EN) {
    foldm::Geom g{F, P, C, E, e0, E - e0 < foldm::EN ? E - e0 : foldm::EN, nw, C * 2 + 16};
    MVF_TRY(foldm_fwd_pass(g, tpw, X, Wq, attn, px, st));
  }
  return MVF_OK;
}
int pool_fold_mma_bwd(int F, int P, int E, int C, const void* X, const float* G, const float* px, const float* attn,
                      float* dWq, cudaStream_t st) {
  int tpw = 0, nw = 0;
  MVF_REQUIRE(foldm::choose_shape(C, &tpw, &nw), MVF_ERR_UNSUPPORTED, "pool_fold_mma: unsupported channel count %d", C);
  for (int e0 = 0; e0 < E; e0 += foldm::EN) {
    foldm::Geom g{F, P, C, E, e0, E - e0 < foldm::EN ? E - e0 : foldm::EN, nw, C * 2 + 16};
    MVF_TRY(foldm_bwd_pass(g, tpw, X, G, px, attn, dWq, st));
  }
  return MVF_OK;
}

}  // namespace mvf

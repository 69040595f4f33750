// Folded entity pooling for bf16 tokens: warp-specialised streaming kernels on the warp-level tensor cores.
//
// Same mathematics and the same single pass over the tokens as the CUDA-core kernels of pool_fold.cu; the contractions run
// as mma.sync.  A first version moved 16-token groups through a 2-slot ring and made all 12 compute warps meet at a named
// barrier once per group, after which every warp redid the softmax bookkeeping: scores -> barrier -> softmax -> pooling
// ran in lock step, the tensor pipe idled during the bookkeeping and only one group (74 KB) was ever in flight per SM
// (measured: 56 % of the HBM copy bandwidth, issue slots 34 % busy, stalls = fixed-latency waits).
//
// Here the ring holds 8-token slots (5 of them at C = 2304: two being consumed, three in flight = 110 KB per SM), and
// three roles communicate through mbarriers only, so nobody waits in lock step:
//
//   producer warp    one bulk async copy per token row into the ring (full / empty mbarriers); the tail group of a frame
//                    loads only its valid rows (the first version re-read 12 rows per frame: 7 % extra HBM traffic)
//   compute warps    own TPW 16-channel tiles each.  Per group n:  S(n+1): partial scores of the NEXT group
//                    S^T[8 ent(+8 pad), 8 tok] += Wq_tile[ent, 16 ch] * X_tile^T[16 ch, 8 tok]       (m16n8k16, A = Wq hi/lo
//                    fragments held in registers, B = X via ldmatrix) -> partial[(n+1)&1][warp] -> arrive pready;
//                    then P(n): wait wready(n), pooling px^T[16 ch, 8 ent] += X_tile^T[16 ch, 8 tok] * w^T[8 tok, 8 ent]
//                    (m16n8k8, A = X via ldmatrix.trans, B = w hi/lo) -> arrive empty(slot)
//   softmax warp     per group: wait pready(n), sum the NW partials (lane = (entity, token)), online softmax over the
//                    frame (running max / sum in registers, log2 domain), write w | rescale factors | 1/sum -> arrive
//                    wready(n); at the end of a frame it also writes the attention rows.  In backward the same warp turns
//                    dA = G X^T into dS = A (dA - delta) with A and delta prefetched one group ahead.
//
// Buffer reuse needs no extra barriers: partial[b] / wbuf[b] of group n are rewritten for group n+2, and every path to that
// write goes through a wait that is only released after the last read of group n (see the ordering notes at the waits).
// Stale ring rows (tail groups) are multiplied by w = 0; the ring is zero-filled once so they are always finite.
#include <math.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace mvf {
namespace foldw {

constexpr int TG = 8;          // tokens per ring slot = tokens per group
constexpr int EN = 8;          // entities per pass (MMA N of the pooling product)
constexpr int WB = 80;         // floats per w buffer: w[8 ent][8 tok] | fc[8] | 1/sum[8]
constexpr int PT = EN * TG;    // floats of one warp's partial tile
constexpr int MAX_SLOTS = 8;

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
      printf("mvf pool_fold_ws: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
// (x, y) -> bf16 hi pair and bf16 lo pair (x = hi + lo up to 2^-17)
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// register budget: 4*TPW accumulator registers per thread (the score fragments live in shared memory); warps are allocated
// in fours, so 16 + 2 warps get the 96 registers of a 640-thread CTA and 12 + 2 warps the 128 of a 512-thread one
__host__ __device__ constexpr int max_compute_warps(int tpw) { return tpw > 9 ? 12 : 16; }
__host__ __device__ constexpr int max_threads(int tpw) { return (max_compute_warps(tpw) + 2) * 32; }

struct Geom {
  int F, P, C;
  int Etot, e0, ne;   // entities in the model, first entity of this pass, entities in this pass (<= 8)
  int NW;             // compute warps
  int pitch;          // shared-memory row pitch in bytes: C*2 + 16 (ldmatrix conflict-free)
  int slots;          // ring slots of TG token rows
};

struct Carve {
  size_t ring, bars, frag, partial, wbuf, table, misc, total;
};
static Carve carve(int C, int NW, int P, int slots, int ne) {
  Carve c;
  c.ring = (size_t)slots * TG * (C * 2 + 16);
  c.bars = 256;                                        // full[8] empty[8] pready[2] wready[2]
  c.frag = (size_t)(C / 16) * ne * 4 * 16;             // A fragments of the scores product: [tile][entity][q] x 16 B
  c.partial = (size_t)2 * NW * PT * 4;
  c.wbuf = (size_t)2 * WB * 4;
  c.table = ((size_t)ne * P * 4 + 15) / 16 * 16;
  c.misc = 64 * 4;
  c.total = c.ring + c.bars + c.frag + c.partial + c.wbuf + c.table + c.misc;
  return c;
}

// ---- shared prologue: carve, zero the ring, barrier init, role split --------------------------------------------------------
struct Sm {
  uint8_t* ring;
  uint64_t *full, *empty, *pready, *wready;
  uint4* frag;
  float *partial, *wbuf, *table, *misc;
};
__device__ __forceinline__ Sm setup(uint8_t* smraw, const Geom& g) {
  Sm s;
  s.ring = smraw;
  const size_t ring_bytes = (size_t)g.slots * TG * g.pitch;
  s.full = reinterpret_cast<uint64_t*>(smraw + ring_bytes);
  s.empty = s.full + MAX_SLOTS;
  s.pready = s.empty + MAX_SLOTS;
  s.wready = s.pready + 2;
  s.frag = reinterpret_cast<uint4*>(smraw + ring_bytes + 256);
  s.partial = reinterpret_cast<float*>(s.frag + (size_t)(g.C / 16) * g.ne * 4);
  s.wbuf = s.partial + 2 * g.NW * PT;
  s.table = s.wbuf + 2 * WB;
  s.misc = s.table + (((size_t)g.ne * g.P + 3) / 4) * 4;
  // zero fill: ring rows that a tail group does not load must hold finite values; wbuf rows of padding entities are read too
  uint4* r4 = reinterpret_cast<uint4*>(smraw);
  for (size_t i = threadIdx.x; i < ring_bytes / 16; i += blockDim.x) r4[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < 2 * WB; i += blockDim.x) s.wbuf[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int k = 0; k < g.slots; ++k) { mbar_init(&s.full[k], 1); mbar_init(&s.empty[k], g.NW); }
    for (int k = 0; k < 2; ++k) { mbar_init(&s.pready[k], g.NW); mbar_init(&s.wready[k], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zero fill before async-proxy bulk copies
  pdl_entry();   // only shared memory was touched so far: this prologue overlaps the previous kernel's tail
  __syncthreads();
  return s;
}

// producer: lane 0 of warp NW streams every group of this CTA's frames through the ring
__device__ __forceinline__ void producer(const Sm& s, const Geom& g, const bf16* __restrict__ X, int nG, int total) {
  const uint32_t row_bytes = (uint32_t)g.C * 2u;
  int slot = 0, use = 0, gi = 0;
  int64_t f = blockIdx.x;
  for (int n = 0; n < total; ++n) {
    if (use > 0) mbar_wait(&s.empty[slot], (use - 1) & 1);
    const int ntok = min(TG, g.P - gi * TG);
    mbar_expect_tx(&s.full[slot], (uint32_t)ntok * row_bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(X) + (f * g.P + (int64_t)gi * TG) * (int64_t)row_bytes;
    uint8_t* dst = s.ring + (size_t)slot * TG * g.pitch;
    for (int r = 0; r < ntok; ++r) bulk_g2s(dst + (size_t)r * g.pitch, src + (int64_t)r * row_bytes, row_bytes, &s.full[slot]);
    if (++slot == g.slots) { slot = 0; ++use; }
    if (++gi == nG) { gi = 0; f += gridDim.x; }
  }
}

// bf16 hi/lo fragments of rows [0, ne) of a [*, C] fp32 matrix for this warp's tiles, kept in SHARED memory: element pairs
// (ch 2q, 2q+1) and (ch 2q+8, 2q+9) of row r8 are the a0 / a2 registers of the m16n8k16 A operand; {hi0, hi1, lo0, lo1} per
// lane.  Only lanes r8 < ne hold data (the padding rows of A are zero), so a register-resident copy would spend 4*TPW
// registers per thread on mostly zeros; 16 B per (tile, entity, q) in shared memory cost one 16-byte load per MMA pair.
template <int TPW>
__device__ __forceinline__ void fill_afrag(uint4* fr, const float* __restrict__ M, int64_t stride, int ne, int ch_base, int q,
                                           int r8, float mul) {
  if (r8 < ne) {
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const float* p = M + (int64_t)r8 * stride + ch_base + t * 16 + 2 * q;
      const float2 v0 = *reinterpret_cast<const float2*>(p);
      const float2 v1 = *reinterpret_cast<const float2*>(p + 8);
      uint4 o;
      split2(v0.x * mul, v0.y * mul, o.x, o.z);
      split2(v1.x * mul, v1.y * mul, o.y, o.w);
      fr[(t * ne + r8) * 4 + q] = o;
    }
  }
  __syncwarp();
}

// S^T[ent, tok] partial over this warp's channels -> pw[ent][tok] (rows r8 < ne).  The entity rows need only 8 of the 16
// rows of the A operand, so the bf16 hi parts sit in rows 0..7 and the lo parts in rows 8..15: ONE MMA per tile yields both
// products (accumulator rows r8 and r8 + 8 live in the same lane and are added at the end).  A row only feeds its own
// output row, so the registers of lanes that own padding rows may hold anything: their results are not stored.
template <int TPW>
__device__ __forceinline__ void scores_phase(uint32_t slot_addr, uint32_t l_off, const uint4* fr, float* pw, int ne, int q,
                                             int r8) {
  float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};   // two chains: even / odd tiles
  // lanes whose A rows are padding (r8 >= ne) skip the load: predicated-off quarter warps issue no shared-memory
  // wavefronts (the fragment loads were half of the kernel's shared-memory traffic), and their registers are don't-care
  const bool loader = r8 < ne;
  const uint4* fl = fr + (loader ? r8 * 4 + q : 0);
  const int tstride = ne * 4;
  uint4 a = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int t = 0; t < TPW; ++t) {
    uint32_t b0, b1;
    ldsm_x2(slot_addr + l_off + t * 32, b0, b1);
    if (loader) a = fl[t * tstride];                   // {hi0, hi1, lo0, lo1}
    if (t & 1) mma16816(sb, a.x, a.z, a.y, a.w, b0, b1);
    else mma16816(sa, a.x, a.z, a.y, a.w, b0, b1);
  }
  if (r8 < ne)
    *reinterpret_cast<float2*>(pw + r8 * TG + 2 * q) = make_float2((sa[0] + sa[2]) + (sb[0] + sb[2]), (sa[1] + sa[3]) + (sb[1] + sb[3]));
}

// acc[t] (= px^T tile [16 ch, 8 columns]) += X^T tile [16 ch, 8 tok] * w^T [8 tok, 8 columns].
// PACK (at most 4 entities in the pass): columns 0..3 carry the bf16 hi parts of w, columns 4..7 the lo parts, so ONE MMA
// per tile does both; the two halves of the accumulator are added when it is written out.  Otherwise the columns are the
// 8 entities and hi / lo take two MMAs.
template <int TPW, bool PACK>
__device__ __forceinline__ void pool_phase(uint32_t slot_addr, uint32_t l_off, uint32_t wh, uint32_t wl, float (&acc)[TPW][4]) {
  if (PACK) {
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      uint32_t a0, a1;
      ldsm_x2_trans(slot_addr + l_off + t * 32, a0, a1);
      mma1688(acc[t], a0, a1, wh);                     // the caller passes the lane's hi-or-lo register in wh
    }
    return;
  }
  // tiles in pairs so that the two dependent MMAs of one accumulator are not issued back to back
#pragma unroll
  for (int t = 0; t + 1 < TPW; t += 2) {
    uint32_t a0, a1, c0, c1;
    ldsm_x2_trans(slot_addr + l_off + t * 32, a0, a1);
    ldsm_x2_trans(slot_addr + l_off + (t + 1) * 32, c0, c1);
    mma1688(acc[t], a0, a1, wl);
    mma1688(acc[t + 1], c0, c1, wl);
    mma1688(acc[t], a0, a1, wh);
    mma1688(acc[t + 1], c0, c1, wh);
  }
  if (TPW & 1) {
    uint32_t a0, a1;
    ldsm_x2_trans(slot_addr + l_off + (TPW - 1) * 32, a0, a1);
    mma1688(acc[TPW - 1], a0, a1, wl);
    mma1688(acc[TPW - 1], a0, a1, wh);
  }
}

// =====================================================================================================================
// forward
// =====================================================================================================================
template <int TPW, bool PACK>
__global__ void __launch_bounds__(max_threads(TPW), 1)
pool_foldw_fwd_kernel(const bf16* __restrict__ X, const float* __restrict__ Wq, float* __restrict__ attn,
                      float* __restrict__ px, const Geom g) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const Sm s = setup(smraw, g);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = g.NW;
  const int nG = (g.P + TG - 1) / TG;
  const int nF = blockIdx.x < g.F ? (g.F - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int total = nF * nG;

  if (warp == NW) {
    if (lane == 0) producer(s, g, X, nG, total);
    return;
  }

  if (warp == NW + 1) {
    // ===== softmax warp: lane = (entity eg [+4], token tok) =====
    const int tok = lane & 7, eg = lane >> 3;
    const bool ev0 = eg < g.ne, ev1 = eg + 4 < g.ne, two = g.ne > 4;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    int gi = 0;
    int64_t f = blockIdx.x;
    float* fin = s.misc;   // [max(8) | 1/sum(8)] of the finished frame
    for (int n = 0; n < total; ++n) {
      const int b = n & 1;
      const int ntok = min(TG, g.P - gi * TG);
      const bool tvalid = tok < ntok, last = gi == nG - 1;
      mbar_wait(&s.pready[b], (n >> 1) & 1);
      const float* pb = s.partial + b * NW * PT + lane;
      float s0 = 0.f, s1 = 0.f;
      for (int w = 0; w < NW; ++w) {
        s0 += pb[w * PT];
        if (two) s1 += pb[w * PT + 32];
      }
      if (ev0 && tvalid) s.table[eg * g.P + gi * TG + tok] = s0;
      if (ev1 && tvalid) s.table[(eg + 4) * g.P + gi * TG + tok] = s1;
      float x0 = (ev0 && tvalid) ? s0 : -INFINITY, x1 = (ev1 && tvalid) ? s1 : -INFINITY;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, o));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
      }
      // every group holds at least one valid token, so the new maxima of real entities are finite
      const float mn0 = fmaxf(m0, x0), mn1 = fmaxf(m1, x1);
      const float w0 = (ev0 && tvalid) ? exp2f(s0 - mn0) : 0.f, w1 = (ev1 && tvalid) ? exp2f(s1 - mn1) : 0.f;
      const float fc0 = ev0 ? exp2f(m0 - mn0) : 1.f, fc1 = ev1 ? exp2f(m1 - mn1) : 1.f;   // first group: 2^-inf = 0
      float t0 = w0, t1 = w1;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        t0 += __shfl_xor_sync(0xffffffffu, t0, o);
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      }
      l0 = l0 * fc0 + t0;
      l1 = l1 * fc1 + t1;
      if (ev0) m0 = mn0;
      if (ev1) m1 = mn1;
      float* wb = s.wbuf + b * WB;
      // wbuf[b] was last read for group n-2; every compute warp finished that read before it arrived on pready(n)
      if (ev0) {
        wb[eg * TG + tok] = w0;
        if (tok == 0) { wb[64 + eg] = fc0; if (last) wb[72 + eg] = 1.f / l0; }
      }
      if (ev1) {
        wb[(eg + 4) * TG + tok] = w1;
        if (tok == 0) { wb[64 + eg + 4] = fc1; if (last) wb[72 + eg + 4] = 1.f / l1; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.wready[b]);
      if (last) {
        // ---- end of frame: attention rows from the raw scores kept in `table` (private to this warp) ----
        if (tok == 0) {
          if (ev0) { fin[eg] = m0; fin[8 + eg] = 1.f / l0; }
          if (ev1) { fin[eg + 4] = m1; fin[8 + eg + 4] = 1.f / l1; }
        }
        __syncwarp();
        for (int i = lane; i < g.ne * g.P; i += 32) {
          const int e = i / g.P, p = i - e * g.P;
          attn[((int64_t)f * g.Etot + g.e0 + e) * g.P + p] = exp2f(s.table[i] - fin[e]) * fin[8 + e];
        }
        __syncwarp();
        m0 = m1 = -INFINITY;
        l0 = l1 = 0.f;
        gi = 0;
        f += gridDim.x;
      } else {
        ++gi;
      }
    }
    return;
  }

  // ===== compute warps =====
  const int q = lane & 3, r8 = lane >> 2;
  const int be = PACK ? (r8 & 3) : r8;                  // entity of this lane's column of the pooling B operand
  const int e0c = PACK ? ((2 * q) & 3) : 2 * q, e1c = e0c + 1;   // entities of this lane's accumulator columns
  const uint32_t ring_u32 = smem_u32(s.ring);
  const uint32_t slot_bytes = (uint32_t)(TG * g.pitch);
  const uint32_t l_off = (uint32_t)((lane & 7) * g.pitch + (warp * TPW * 16 + ((lane >> 3) & 1) * 8) * 2);
  // the scores live in the log2 domain (log2 e folded into the Wq fragments)
  uint4* fr = s.frag + (size_t)warp * TPW * g.ne * 4;
  fill_afrag<TPW>(fr, Wq + (int64_t)g.e0 * g.C, g.C, g.ne, warp * TPW * 16, q, r8, 1.4426950408889634f);
  float acc[TPW][4];
#pragma unroll
  for (int t = 0; t < TPW; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

  int s_slot = 0, s_use = 0;   // ring position of the next group to score
  int p_slot = 0;              // ring position of the group being pooled
  int gi = 0;
  int64_t f = blockIdx.x;
  auto score_next = [&](int n1) {
    mbar_wait(&s.full[s_slot], s_use & 1);
    // partial[(n1)&1] was last read for group n1-2; that read finished before wready(n1-2) which this warp has waited on
    scores_phase<TPW>(ring_u32 + s_slot * slot_bytes, l_off, fr, s.partial + ((n1 & 1) * NW + warp) * PT, g.ne, q, r8);
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.pready[n1 & 1]);
    if (++s_slot == g.slots) { s_slot = 0; ++s_use; }
  };
  if (total > 0) score_next(0);
  for (int n = 0; n < total; ++n) {
    if (n + 1 < total) score_next(n + 1);
    const int b = n & 1;
    const bool last = gi == nG - 1;
    mbar_wait(&s.wready[b], (n >> 1) & 1);
    const float* wb = s.wbuf + b * WB;
    const float2 w = *reinterpret_cast<const float2*>(wb + be * TG + 2 * q);   // rows of padding entities stay zero
    const float f0 = (e0c < g.ne) ? wb[64 + e0c] : 1.f;
    const float f1 = (e1c < g.ne) ? wb[64 + e1c] : 1.f;
    float i0 = 0.f, i1 = 0.f;
    if (last) { i0 = wb[72 + e0c]; i1 = wb[72 + e1c]; }
    uint32_t wh, wl;
    split2(w.x, w.y, wh, wl);
    if (PACK && r8 >= 4) wh = wl;
    // rescale the accumulators when a running maximum moved (columns = entities 2q, 2q+1)
    if (__any_sync(0xffffffffu, f0 != 1.f || f1 != 1.f)) {
#pragma unroll
      for (int t = 0; t < TPW; ++t) { acc[t][0] *= f0; acc[t][1] *= f1; acc[t][2] *= f0; acc[t][3] *= f1; }
    }
    pool_phase<TPW, PACK>(ring_u32 + p_slot * slot_bytes, l_off, wh, wl, acc);
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.empty[p_slot]);
    if (++p_slot == g.slots) p_slot = 0;
    if (last) {
      float* base = px + ((int64_t)f * g.Etot + g.e0) * g.C + warp * TPW * 16;
      const bool writer = !PACK || q < 2;
#pragma unroll
      for (int t = 0; t < TPW; ++t) {
        float v0 = acc[t][0], v1 = acc[t][1], v2 = acc[t][2], v3 = acc[t][3];
        if (PACK) {   // hi half (lanes q < 2) + lo half (lanes q >= 2) of the same entity columns
          v0 += __shfl_xor_sync(0xffffffffu, v0, 2); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
          v2 += __shfl_xor_sync(0xffffffffu, v2, 2); v3 += __shfl_xor_sync(0xffffffffu, v3, 2);
        }
        const int ch = t * 16 + r8;
        if (writer && e0c < g.ne) { base[(int64_t)e0c * g.C + ch] = v0 * i0; base[(int64_t)e0c * g.C + ch + 8] = v2 * i0; }
        if (writer && e1c < g.ne) { base[(int64_t)e1c * g.C + ch] = v1 * i1; base[(int64_t)e1c * g.C + ch + 8] = v3 * i1; }
        acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
      }
      gi = 0;
      f += gridDim.x;
    } else {
      ++gi;
    }
  }
}

// =====================================================================================================================
// backward:  dWq[e, :] += sum_frames sum_p A[e,p] (G_e . x_p - delta_e) x_p,   delta_e = G_e . px_e
// =====================================================================================================================
template <int TPW, bool PACK>
__global__ void __launch_bounds__(max_threads(TPW), 1)
pool_foldw_bwd_kernel(const bf16* __restrict__ X, const float* __restrict__ G, const float* __restrict__ px,
                      const float* __restrict__ attn, const float* __restrict__ delta, float* __restrict__ dWq, const Geom g) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const Sm s = setup(smraw, g);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = g.NW;
  const int nG = (g.P + TG - 1) / TG;
  const int nF = blockIdx.x < g.F ? (g.F - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int total = nF * nG;

  if (warp == NW) {
    if (lane == 0) producer(s, g, X, nG, total);
    return;
  }

  if (warp == NW + 1) {
    // ===== dS warp: lane = (entity eg [+4], token tok); A and delta of the next group are fetched one group ahead =====
    const int tok = lane & 7, eg = lane >> 3;
    const bool ev0 = eg < g.ne, ev1 = eg + 4 < g.ne, two = g.ne > 4;
    auto frame_delta = [&](int64_t f, int e) -> float {
      if (delta) return delta[f * g.Etot + g.e0 + e];
      return 0.f;   // filled by the warp-wide reduction below when no delta array is given
    };
    auto fetch = [&](int64_t f, int gi, float& a0, float& a1, float& d0, float& d1) {
      const int p = gi * TG + tok;
      const bool tv = p < g.P;
      a0 = (ev0 && tv) ? attn[((int64_t)f * g.Etot + g.e0 + eg) * g.P + p] : 0.f;
      a1 = (ev1 && tv) ? attn[((int64_t)f * g.Etot + g.e0 + eg + 4) * g.P + p] : 0.f;
      d0 = ev0 ? frame_delta(f, eg) : 0.f;
      d1 = ev1 ? frame_delta(f, eg + 4) : 0.f;
    };
    // without a precomputed delta (stand-alone C-ABI call) the warp reduces <G_e, px_e> at the start of every frame
    auto reduce_delta = [&](int64_t f) {
      for (int e = 0; e < g.ne; ++e) {
        const float4* gp = reinterpret_cast<const float4*>(G + ((int64_t)f * g.Etot + g.e0 + e) * g.C);
        const float4* pp = reinterpret_cast<const float4*>(px + ((int64_t)f * g.Etot + g.e0 + e) * g.C);
        float acc = 0.f;
        for (int i = lane; i < g.C / 4; i += 32) {
          const float4 a = gp[i], c = pp[i];
          acc += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s.misc[e] = acc;
      }
      __syncwarp();
    };
    int gi = 0;
    int64_t f = blockIdx.x;
    float a0 = 0.f, a1 = 0.f, d0 = 0.f, d1 = 0.f;
    if (total > 0) {
      fetch(f, 0, a0, a1, d0, d1);
      if (!delta) { reduce_delta(f); d0 = ev0 ? s.misc[eg] : 0.f; d1 = ev1 ? s.misc[eg + 4] : 0.f; __syncwarp(); }
    }
    for (int n = 0; n < total; ++n) {
      const int b = n & 1;
      int gi_n = gi + 1;
      int64_t f_n = f;
      if (gi_n == nG) { gi_n = 0; f_n += gridDim.x; }
      float na0 = 0.f, na1 = 0.f, nd0 = 0.f, nd1 = 0.f;
      if (n + 1 < total) {
        fetch(f_n, gi_n, na0, na1, nd0, nd1);
        if (!delta) {
          if (gi_n == 0) { reduce_delta(f_n); nd0 = ev0 ? s.misc[eg] : 0.f; nd1 = ev1 ? s.misc[eg + 4] : 0.f; __syncwarp(); }
          else { nd0 = d0; nd1 = d1; }
        }
      }
      mbar_wait(&s.pready[b], (n >> 1) & 1);
      const float* pb = s.partial + b * NW * PT + lane;
      float s0 = 0.f, s1 = 0.f;
      for (int w = 0; w < NW; ++w) {
        s0 += pb[w * PT];
        if (two) s1 += pb[w * PT + 32];
      }
      float* wb = s.wbuf + b * WB;
      if (ev0) wb[eg * TG + tok] = a0 * (s0 - d0);          // a = 0 for tokens past the end of the frame
      if (ev1) wb[(eg + 4) * TG + tok] = a1 * (s1 - d1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.wready[b]);
      a0 = na0; a1 = na1; d0 = nd0; d1 = nd1;
      gi = gi_n;
      f = f_n;
    }
    return;
  }

  // ===== compute warps =====
  const int q = lane & 3, r8 = lane >> 2;
  const int be = PACK ? (r8 & 3) : r8;                  // entity of this lane's column of the pooling B operand
  const int e0c = PACK ? ((2 * q) & 3) : 2 * q, e1c = e0c + 1;   // entities of this lane's accumulator columns
  const uint32_t ring_u32 = smem_u32(s.ring);
  const uint32_t slot_bytes = (uint32_t)(TG * g.pitch);
  const uint32_t l_off = (uint32_t)((lane & 7) * g.pitch + (warp * TPW * 16 + ((lane >> 3) & 1) * 8) * 2);
  uint4* fr = s.frag + (size_t)warp * TPW * g.ne * 4;
  float acc[TPW][4];
#pragma unroll
  for (int t = 0; t < TPW; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

  int s_slot = 0, s_use = 0, s_gi = 0;
  int64_t s_f = blockIdx.x;    // frame of the next group to score: its G rows are the A operand
  int p_slot = 0;
  auto score_next = [&](int n1) {
    if (s_gi == 0) fill_afrag<TPW>(fr, G + ((int64_t)s_f * g.Etot + g.e0) * g.C, g.C, g.ne, warp * TPW * 16, q, r8, 1.f);
    mbar_wait(&s.full[s_slot], s_use & 1);
    scores_phase<TPW>(ring_u32 + s_slot * slot_bytes, l_off, fr, s.partial + ((n1 & 1) * NW + warp) * PT, g.ne, q, r8);
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.pready[n1 & 1]);
    if (++s_slot == g.slots) { s_slot = 0; ++s_use; }
    if (++s_gi == nG) { s_gi = 0; s_f += gridDim.x; }
  };
  if (total > 0) score_next(0);
  for (int n = 0; n < total; ++n) {
    if (n + 1 < total) score_next(n + 1);
    const int b = n & 1;
    mbar_wait(&s.wready[b], (n >> 1) & 1);
    const float2 w = *reinterpret_cast<const float2*>(s.wbuf + b * WB + be * TG + 2 * q);
    uint32_t wh, wl;
    split2(w.x, w.y, wh, wl);
    if (PACK && r8 >= 4) wh = wl;
    pool_phase<TPW, PACK>(ring_u32 + p_slot * slot_bytes, l_off, wh, wl, acc);
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.empty[p_slot]);
    if (++p_slot == g.slots) p_slot = 0;
  }
  if (nF > 0) {
    float* base = dWq + (int64_t)g.e0 * g.C + warp * TPW * 16;
    const bool writer = !PACK || q < 2;
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      float v0 = acc[t][0], v1 = acc[t][1], v2 = acc[t][2], v3 = acc[t][3];
      if (PACK) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, 2); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
        v2 += __shfl_xor_sync(0xffffffffu, v2, 2); v3 += __shfl_xor_sync(0xffffffffu, v3, 2);
      }
      const int ch = t * 16 + r8;
      if (writer && e0c < g.ne) { atomicAdd(base + (int64_t)e0c * g.C + ch, v0); atomicAdd(base + (int64_t)e0c * g.C + ch + 8, v2); }
      if (writer && e1c < g.ne) { atomicAdd(base + (int64_t)e1c * g.C + ch, v1); atomicAdd(base + (int64_t)e1c * g.C + ch + 8, v3); }
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

// tiles per warp / compute warps for C channels (C % 16 == 0): prefer a warp count that is a multiple of 4, then more warps
static bool choose_shape(int C, int P, int ne, int* tpw_out, int* nw_out, int* slots_out) {
  const int CT = C / 16;
  static const int cand[] = {12, 9, 8, 6, 4, 3, 2, 1};
  int best_tpw = 0, best_nw = 0, best_score = -1;
  for (int tpw : cand) {
    if (CT % tpw) continue;
    const int nw = CT / tpw;
    if (nw > max_compute_warps(tpw)) continue;
    const int score = (nw % 4 == 0 ? 100 : 0) + nw;
    if (score > best_score) { best_score = score; best_tpw = tpw; best_nw = nw; }
  }
  static int forced_tpw = -1;   // MVF_FOLD_TPW: tiles per compute warp override for A/B measurements
  if (forced_tpw < 0) {
    const char* e = getenv("MVF_FOLD_TPW");
    forced_tpw = e ? atoi(e) : 0;
  }
  if (forced_tpw > 0 && CT % forced_tpw == 0 && CT / forced_tpw <= max_compute_warps(forced_tpw)) {
    bool known = false;
    for (int tpw : cand) known = known || tpw == forced_tpw;
    if (known) { best_tpw = forced_tpw; best_nw = CT / forced_tpw; best_score = 0; }
  }
  if (best_score < 0) return false;
  int slots = MAX_SLOTS;
  while (slots >= 3 && carve(C, best_nw, P, slots, ne).total > 227 * 1024) --slots;
  if (slots < 3) return false;
  static int forced = -1;   // MVF_FOLD_SLOTS: ring depth override for A/B measurements
  if (forced < 0) {
    const char* e = getenv("MVF_FOLD_SLOTS");
    forced = e ? atoi(e) : 0;
  }
  if (forced >= 3 && forced <= slots) slots = forced;
  *tpw_out = best_tpw;
  *nw_out = best_nw;
  *slots_out = slots;
  return true;
}

static int g_reserve_sms = 0;   // SMs left free by the backward kernel (pool_fold_reserve_sms)

struct Plan {
  int F = -1, P = -1, C = -1, ne = -1, slots = -1, reserve = -1;
  int grid = 0;
  size_t smem = 0;
};

template <typename KernelT>
static int plan(KernelT kernel, Plan& pl, const Geom& g, int reserve = 0) {
  if (pl.F == g.F && pl.P == g.P && pl.C == g.C && pl.ne == g.ne && pl.slots == g.slots && pl.reserve == reserve) return MVF_OK;
  Carve cv = carve(g.C, g.NW, g.P, g.slots, g.ne);
  MVF_REQUIRE(cv.total <= 227 * 1024, MVF_ERR_UNSUPPORTED, "pool_fold_ws: %d channels need %zu B of shared memory", g.C, cv.total);
  MVF_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cv.total));
  int sms = num_sms() - reserve;
  if (sms < 1) sms = 1;
  pl.grid = g.F < sms ? g.F : sms;     // one persistent CTA per SM (the ring takes most of the shared memory)
  pl.reserve = reserve;
  pl.smem = cv.total;
  pl.F = g.F; pl.P = g.P; pl.C = g.C; pl.ne = g.ne; pl.slots = g.slots;
  return MVF_OK;
}

template <int TPW>
static int fwd_launch(const Geom& g, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  if (g.ne <= 4) {
    static thread_local Plan pl;
    MVF_TRY(plan(pool_foldw_fwd_kernel<TPW, true>, pl, g));
    launch_k(pool_foldw_fwd_kernel<TPW, true>, pl.grid, (g.NW + 2) * 32, pl.smem, st, (const bf16*)X, Wq, attn, px, g);
  } else {
    static thread_local Plan pl;
    MVF_TRY(plan(pool_foldw_fwd_kernel<TPW, false>, pl, g));
    launch_k(pool_foldw_fwd_kernel<TPW, false>, pl.grid, (g.NW + 2) * 32, pl.smem, st, (const bf16*)X, Wq, attn, px, g);
  }
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
template <int TPW>
static int bwd_launch(const Geom& g, const void* X, const float* G, const float* px, const float* attn, const float* delta,
                      float* dWq, cudaStream_t st) {
  if (g.ne <= 4) {
    static thread_local Plan pl;
    MVF_TRY(plan(pool_foldw_bwd_kernel<TPW, true>, pl, g, g_reserve_sms));
    launch_k(pool_foldw_bwd_kernel<TPW, true>, pl.grid, (g.NW + 2) * 32, pl.smem, st, (const bf16*)X, G, px, attn, delta, dWq, g);
  } else {
    static thread_local Plan pl;
    MVF_TRY(plan(pool_foldw_bwd_kernel<TPW, false>, pl, g, g_reserve_sms));
    launch_k(pool_foldw_bwd_kernel<TPW, false>, pl.grid, (g.NW + 2) * 32, pl.smem, st, (const bf16*)X, G, px, attn, delta, dWq, g);
  }
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace foldw

int pool_fold_reserve_sms(int n) {
  const int prev = foldw::g_reserve_sms;
  foldw::g_reserve_sms = n < 0 ? 0 : n;
  return prev;
}

bool pool_fold_ws_supported(int dtype, int C, int P) {
  int tpw, nw, slots;
  if (dtype != MVF_BF16 || C <= 0 || C % 16 != 0 || P < 1) return false;
  return foldw::choose_shape(C, P, foldw::EN, &tpw, &nw, &slots);
}

#define MVF_FOLDW_DISPATCH(FN, ...)                      \
  switch (tpw) {                                         \
    case 1: return foldw::FN<1>(__VA_ARGS__);            \
    case 2: return foldw::FN<2>(__VA_ARGS__);            \
    case 3: return foldw::FN<3>(__VA_ARGS__);            \
    case 4: return foldw::FN<4>(__VA_ARGS__);            \
    case 6: return foldw::FN<6>(__VA_ARGS__);            \
    case 8: return foldw::FN<8>(__VA_ARGS__);            \
    case 9: return foldw::FN<9>(__VA_ARGS__);            \
    default: return foldw::FN<12>(__VA_ARGS__);          \
  }

static int foldw_fwd_pass(const foldw::Geom& g, int tpw, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  MVF_FOLDW_DISPATCH(fwd_launch, g, X, Wq, attn, px, st)
}
static int foldw_bwd_pass(const foldw::Geom& g, int tpw, const void* X, const float* G, const float* px, const float* attn,
                          const float* delta, float* dWq, cudaStream_t st) {
  MVF_FOLDW_DISPATCH(bwd_launch, g, X, G, px, attn, delta, dWq, st)
}

int pool_fold_ws_fwd(int F, int P, int E, int C, const void* X, const float* Wq, float* attn, float* px, cudaStream_t st) {
  for (int e0 = 0; e0 < E; e0 += foldw::EN) {
    const int ne = E - e0 < foldw::EN ? E - e0 : foldw::EN;
    int tpw = 0, nw = 0, slots = 0;
    MVF_REQUIRE(foldw::choose_shape(C, P, ne, &tpw, &nw, &slots), MVF_ERR_UNSUPPORTED, "pool_fold_ws: unsupported channel count %d", C);
    foldw::Geom g{F, P, C, E, e0, ne, nw, C * 2 + 16, slots};
    MVF_TRY(foldw_fwd_pass(g, tpw, X, Wq, attn, px, st));
  }
  return MVF_OK;
}
// delta [F*E] = <G_row, px_row> may be null (the kernel then reduces it per frame, slower)
int pool_fold_ws_bwd(int F, int P, int E, int C, const void* X, const float* G, const float* px, const float* attn,
                     const float* delta, float* dWq, cudaStream_t st) {
  for (int e0 = 0; e0 < E; e0 += foldw::EN) {
    const int ne = E - e0 < foldw::EN ? E - e0 : foldw::EN;
    int tpw = 0, nw = 0, slots = 0;
    MVF_REQUIRE(foldw::choose_shape(C, P, ne, &tpw, &nw, &slots), MVF_ERR_UNSUPPORTED, "pool_fold_ws: unsupported channel count %d", C);
    foldw::Geom g{F, P, C, E, e0, ne, nw, C * 2 + 16, slots};
    MVF_TRY(foldw_bwd_pass(g, tpw, X, G, px, attn, delta, dWq, st));
  }
  return MVF_OK;
}

}  // namespace mvf

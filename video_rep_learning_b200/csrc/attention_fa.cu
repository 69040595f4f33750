// Temporal self-attention for long sequences (S > 64, d_k = 32) on the 5th-generation tensor cores: flash-attention style
// forward and backward with tcgen05.mma (accumulators in TMEM), operands staged by 3-D TMA tensor maps, online softmax in
// the exp2 domain.  models/utils.py:11-44 + 47-108 with the [B,1,1,S] key mask; S = E*T is 480 at BASELINE cfg4, 3840 at
// cfg5 and up to ~12 k in whole-video evaluation (evaluate.py:45-62), where the reference materialises [B,8,S,S] scores.
//
// Precision: the chain behind the pooling feeds SCL's 1/temperature, so every contraction uses bf16 hi/lo operand splits
// (x = hi + lo, fp32 accumulation in TMEM, ~16 mantissa bits) like the forward GEMMs (gemm_tc.cu SPLIT3) and the S <= 64
// kernels (attention_tc.cu).  A pre-pass writes each operand once per layer "row-split":
//     RS_x [slab][S][64] bf16:  columns 0..31 = hi(x[s, :]),  32..63 = lo(x[s, :])          slab = view * heads + head
// One 128-byte row is one SWIZZLE_128B atom row, and the SAME shared-memory tile serves both kinds of contraction:
//   * contraction over d_k (S = Q K^T, dP = dO V^T, ...): A and B K-major; hi*hi + hi*lo + lo*hi are the same two
//     descriptors advanced by 0 / 64 bytes inside the row (6 MMAs of K = 16);
//   * contraction over tokens (O = P V, dQ = dS K, dV = P^T dO, dK = dS^T Q): the tile is the B operand in MN-major form
//     (tokens are the K rows, the 64 columns hi|lo are N), so ONE product gives D[:, 0:32] = A x_hi and D[:, 32:64] = A x_lo;
//     A = the probabilities (hi and lo tiles written to shared memory by the softmax warps), and the two halves of the
//     accumulator are added when it is read out.  No transposed copies of Q / K / V / dO exist anywhere.
//
//   forward   CTA = 128 queries x one slab; loop over 64-key tiles:  S = Q K^T (TMEM, double buffered) -> softmax warps
//             (one thread per row, tcgen05.ld) -> P hi|lo into shared memory -> PV (TMEM) -> O kept in registers
//             (O = O * alpha + PV).  2 CTAs per SM, 3-deep K / V rings.
//   dQ        CTA = 128 queries; per key tile S = Q K^T and dP = dO V^T (TMEM) -> dS = P (dP - delta) / sqrt(dk) hi|lo
//             into shared memory -> dQ += dS K accumulated in TMEM over the whole loop.  2 CTAs per SM.
//   dK, dV    CTA = 128 keys; per 64-query tile S^T = K Q^T, dP^T = V dO^T (TMEM, double buffered) -> P^T, dS^T hi|lo
//             into shared memory (two softmax warpgroups, 32 columns each) -> dV += P^T dO, dK += dS^T Q in TMEM; 4-deep ring.
// Every score is recomputed in the orientation its consumer needs, so no transposed accumulator travels through memory
// (the scheme of attention_tc.cu, now streaming).  d_k = 32 makes the kernels exp/ALU-bound, not tensor-bound
// (128 tensor FLOPs per exp): the roofline reported for them is MUFU ex2 / issue throughput.
#include <math.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace mvf {
namespace fa {

using namespace tc;

constexpr int BQ = 128;     // rows per CTA (UMMA M)
constexpr int BK = 64;      // columns per inner tile
constexpr int DK = 32;
constexpr int ROW = 128;    // bytes per operand row (64 bf16)
constexpr int T64 = 64 * ROW;    // a 64-row tile
constexpr int T128 = 128 * ROW;  // a 128-row tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr int MAX_MASK_WORDS = 512;    // S <= 16384

__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split8(const float* v, uint4& h, uint4& l) {
  split2(v[0], v[1], h.x, l.x);
  split2(v[2], v[3], h.y, l.y);
  split2(v[4], v[5], h.z, l.z);
  split2(v[6], v[7], h.w, l.w);
}
// 2^x on the MUFU (ex2.approx.ftz: 2^-inf = +0); exp2f() wraps the same instruction in range fix-ups the kernels do not need
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// 16-byte chunk `chunk` of row r of a SWIZZLE_128B K-major tile: hi and lo tiles
__device__ __forceinline__ void store_chunk(uint32_t tile_hi, uint32_t tile_lo, int r, int chunk, const uint4& h, const uint4& l) {
  const uint32_t off = (uint32_t)(r * ROW + ((chunk ^ (r & 7)) << 4));
  sts128(tile_hi + off, h);
  sts128(tile_lo + off, l);
}
// the view's key-mask words -> shared memory (called by the `n` softmax threads, ids 0..n-1, after pdl_entry)
__device__ __forceinline__ void load_mask_words(uint32_t* smask, const uint32_t* __restrict__ mb, int words, int tid, int n) {
  for (int i = tid; i < words + 2; i += n) smask[i] = i < words ? __ldg(mb + i) : 0u;
}
// 1-D bulk copy global -> shared, completion on an mbarrier (bytes % 16 == 0, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// pre-pass: fp32 [B*S, ld] (head h at columns h*32 of each source) -> row-split bf16 hi|lo operands (+ key-mask bit words)
// ---------------------------------------------------------------------------------------------------------------------
struct PrepSrc {
  const float* src;   // first row of view 0, first column of head 0
  int64_t ld;         // row stride (floats)
  bf16* rs;           // [slab][S][64]
};
struct PrepArgs {
  PrepSrc m[4];
  int n;
  const float* keymask;   // [B, S] or null
  uint32_t* maskbits;     // [B][mask_words]: bit j of word w = key 32 w + j may be attended
  int mask_words;
};

__global__ void __launch_bounds__(256) fa_prep_kernel(const PrepArgs a, int S, int heads) {
  pdl_entry();
  const int b = blockIdx.z, h = blockIdx.y, s0 = blockIdx.x * 32, tid = threadIdx.x;
  const int64_t slab = (int64_t)b * heads + h;
  const int r = tid >> 3, c4 = tid & 7;
  const int s = s0 + r;
  if (h == 0 && tid < 32) {
    const int j = s0 + tid;
    const bool ok = j < S && (a.keymask == nullptr || a.keymask[(int64_t)b * S + j] != 0.f);
    const uint32_t w = __ballot_sync(0xffffffffu, ok);
    if (tid == 0) a.maskbits[(int64_t)b * a.mask_words + blockIdx.x] = w;
  }
  if (s >= S) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i >= a.n) break;
    const PrepSrc& m = a.m[i];
    const float4 v = *reinterpret_cast<const float4*>(m.src + ((int64_t)b * S + s) * m.ld + h * DK + 4 * c4);
    uint2 hi, lo;
    split2(v.x, v.y, hi.x, lo.x);
    split2(v.z, v.w, hi.y, lo.y);
    bf16* row = m.rs + (slab * S + s) * 64;
    *reinterpret_cast<uint2*>(row + 4 * c4) = hi;
    *reinterpret_cast<uint2*>(row + 32 + 4 * c4) = lo;
  }
}

// backward row statistics: L2[slab][Sq] = -lse * log2(e) (-inf where the row has no valid key or is padding, so that
// 2^(s + L2) = 0) and D[slab][Sq] = <dO_i, O_i> / sqrt(dk) over the head's channels
__global__ void __launch_bounds__(256) fa_stats_kernel(const float* __restrict__ ctx, const float* __restrict__ d_ctx,
                                                       const float* __restrict__ lse, float* __restrict__ L2,
                                                       float* __restrict__ Dl, int S, int Sq, int heads, int H) {
  pdl_entry();
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * 32 + (threadIdx.x >> 3), c4 = threadIdx.x & 7;
  const int64_t slab = (int64_t)b * heads + h;
  float dl = 0.f;
  if (i < S) {
    const int64_t off = ((int64_t)b * S + i) * H + h * DK + 4 * c4;
    const float4 g = *reinterpret_cast<const float4*>(d_ctx + off);
    const float4 o = *reinterpret_cast<const float4*>(ctx + off);
    dl = g.x * o.x + g.y * o.y + g.z * o.z + g.w * o.w;
  }
  dl += __shfl_xor_sync(0xffffffffu, dl, 1);
  dl += __shfl_xor_sync(0xffffffffu, dl, 2);
  dl += __shfl_xor_sync(0xffffffffu, dl, 4);
  if (c4 == 0 && i < Sq) {
    float l2 = -INFINITY;
    if (i < S) {
      const float lv = lse[slab * S + i];
      l2 = (lv == -INFINITY) ? -INFINITY : -lv * LOG2E;
    }
    L2[slab * Sq + i] = l2;
    Dl[slab * Sq + i] = dl * rsqrtf((float)DK);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// the two MMA patterns
// ---------------------------------------------------------------------------------------------------------------------
// D[128 x N] = A[128 x 32] * B[N x 32]^T on row-split operands (contraction over d_k): hi*hi + hi*lo + lo*hi
__device__ __forceinline__ void mma_rowsplit(uint32_t tmem_d, uint32_t sa, uint32_t sb, uint32_t idesc, uint32_t leader) {
  const uint64_t da = make_smem_desc(sa, true, 0, 2), db = make_smem_desc(sb, true, 0, 2);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint64_t hi = (uint64_t)(2 * j), lo = (uint64_t)(4 + 2 * j);   // 16-byte units inside the 128-byte row
    umma_f16_p(tmem_d, da + lo, db + hi, idesc, j > 0 ? 1u : 0u, leader);   // small terms first
    umma_f16_p(tmem_d, da + hi, db + lo, idesc, 1u, leader);
    umma_f16_p(tmem_d, da + hi, db + hi, idesc, 1u, leader);
  }
}
// D[128 x 64] (+)= A[128 x 64 tokens] * X[64 tokens x (hi 32 | lo 32)]: A = hi and lo probability tiles (K-major), X = a
// row-split tile used as an MN-major B operand (token rows are the K rows); columns 0..31 of D collect A x_hi, 32..63 A x_lo
__device__ __forceinline__ void mma_tokens(uint32_t tmem_d, uint32_t sa_hi, uint32_t sa_lo, uint32_t sb, uint32_t idesc, bool accumulate,
                                           uint32_t leader) {
  const uint64_t dah = make_smem_desc(sa_hi, true, 0, 2), dal = make_smem_desc(sa_lo, true, 0, 2);
  const uint64_t db = make_smem_desc(sb, false, 64, 2);
#pragma unroll
  for (int k = 0; k < 4; ++k) {                                          // 4 x 16 tokens
    const uint64_t oa = (uint64_t)(2 * k), ob = (uint64_t)(128 * k);      // 32 bytes along an A row; 16 token rows of B
    umma_f16_p(tmem_d, dal + oa, db + ob, idesc, (accumulate || k > 0) ? 1u : 0u, leader);
    umma_f16_p(tmem_d, dah + oa, db + ob, idesc, 1u, leader);
  }
}

// the same product with the probabilities in TENSOR MEMORY (A operand of the "TS" MMA form): ta_hi / ta_lo = TMEM column of
// the packed bf16 pairs (32 columns for 64 tokens; 8 columns per K = 16 step), written there by the softmax warps with
// tcgen05.st -- no shared-memory store by 128 threads, no shared-memory operand read by the tensor core for A
__device__ __forceinline__ void mma_tokens_ts(uint32_t tmem_d, uint32_t ta_hi, uint32_t ta_lo, uint32_t sb, uint32_t idesc, bool accumulate,
                                              uint32_t leader) {
  const uint64_t db = make_smem_desc(sb, false, 64, 2);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t ob = (uint64_t)(128 * k);
    umma_ts_f16_p(tmem_d, ta_lo + 8 * k, db + ob, idesc, (accumulate || k > 0) ? 1u : 0u, leader);
    umma_ts_f16_p(tmem_d, ta_hi + 8 * k, db + ob, idesc, 1u, leader);
  }
}
// eight fp32 values -> four packed bf16 pairs of the hi part and of the lo part
__device__ __forceinline__ void split8u(const float* v, uint32_t* h, uint32_t* l) {
  split2(v[0], v[1], h[0], l[0]);
  split2(v[2], v[3], h[1], l[1]);
  split2(v[4], v[5], h[2], l[2]);
  split2(v[6], v[7], h[3], l[3]);
}

struct Params {
  int S, Sq, heads, H;
  int mask_words;
  const uint32_t* maskbits;
  const float* L2;     // [slab][Sq]
  const float* Dl;     // [slab][Sq]
  float* ctx;          // forward output [B*S, H]
  float* lse;          // forward output [slab][S]
  float* d_qkv;        // backward output [B*S, 3H]
  long long* dbg;      // MVF_FA_DBG=1: clock64() stamps of CTA 0 of the dK/dV kernel (debug aid, null otherwise)
};
#define FA_STAMP(slot, j) do { if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 24 && (threadIdx.x & 31) == 0) p.dbg[(j) * 16 + (slot)] = clock64(); } while (0)

// =====================================================================================================================
// forward
// =====================================================================================================================
constexpr int FWD_THREADS = 192;   // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 softmax (TMEM lane quadrant = warp % 4)
constexpr int FWD_NK = 3, FWD_NV = 3;
constexpr int FWD_SMEM = T128 + (FWD_NK + FWD_NV) * T64 + 2 * T128;   // Q | K ring | V ring | P hi, lo

template <bool ATM>   // ATM: probabilities handed to the tensor core through tensor memory instead of shared memory
__global__ void __launch_bounds__(FWD_THREADS, 2)
fa_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
              const __grid_constant__ CUtensorMap map_v, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t smask[MAX_MASK_WORDS + 2];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + T128;
  uint8_t* sV = sK + FWD_NK * T64;
  uint8_t* sPh = sV + FWD_NV * T64;
  uint8_t* sPl = sPh + T128;
  uint64_t* bars = (uint64_t*)(sPl + T128);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;              // [3]
  uint64_t* k_empty = k_full + FWD_NK;      // [3]
  uint64_t* v_full = k_empty + FWD_NK;      // [3]
  uint64_t* v_empty = v_full + FWD_NV;      // [3]
  uint64_t* s_full = v_empty + FWD_NV;      // [2]
  uint64_t* s_empty = s_full + 2;           // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = (uint32_t*)(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int slab = b * p.heads + h;
  const int nt = (p.S + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FWD_NK; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < FWD_NV; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // columns: S0 [0,64) | S1 [64,128) | PV [128,192) | P hi [192,224) | P lo [224,256)
  pdl_entry();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, T128);
      tma_load_3d(&map_q, q_full, sQ, 0, q0, slab);
      for (int j = 0; j < nt; ++j) {
        const int sk = j % FWD_NK, sv = j % FWD_NV;
        mbar_wait(&k_empty[sk], ((j / FWD_NK) & 1) ^ 1, 10);
        mbar_expect_tx(&k_full[sk], T64);
        tma_load_3d(&map_k, &k_full[sk], sK + sk * T64, 0, j * BK, slab);
        mbar_wait(&v_empty[sv], ((j / FWD_NV) & 1) ^ 1, 11);
        mbar_expect_tx(&v_full[sv], T64);
        tma_load_3d(&map_v, &v_full[sv], sV + sv * T64, 0, j * BK, slab);
      }
    }
  } else if (warp == 1) {
    // every lane runs the protocol with warp-uniform operands; the elected lane issues (see umma_f16_p)
    {
      const uint32_t leader = elect_leader();
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0), sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t aQ = sb, aK = sb + T128, aV = aK + FWD_NK * T64, aPh = aV + FWD_NV * T64, aPl = aPh + T128;
      const uint32_t idesc_s = make_idesc(BK, true, true, 2), idesc_o = make_idesc(64, true, false, 2);
      auto issue_s = [&](int j) {
        const int st = j & 1, sk = j % FWD_NK;
        mbar_wait(&k_full[sk], (j / FWD_NK) & 1, 12);
        mbar_wait(&s_empty[st], ((j >> 1) & 1) ^ 1, 13);
        tcgen05_fence_after();
        mma_rowsplit(tb + st * BK, aQ, aK + sk * T64, idesc_s, leader);
        umma_commit_p(&k_empty[sk], leader);
        umma_commit_p(&s_full[st], leader);
      };
      mbar_wait(q_full, 0, 14);
      issue_s(0);
      if (nt > 1) issue_s(1);
      for (int j = 0; j < nt; ++j) {
        // two tiles ahead: S_{j+2} reuses the buffer of S_j, free as soon as the softmax warps have loaded tile j
        if (j + 2 < nt) issue_s(j + 2);
        const int sv = j % FWD_NV;
        mbar_wait(p_full, j & 1, 15);
        mbar_wait(&v_full[sv], (j / FWD_NV) & 1, 16);
        tcgen05_fence_after();
        if constexpr (ATM) mma_tokens_ts(tb + 2 * BK, tb + 192, tb + 224, aV + sv * T64, idesc_o, false, leader);
        else mma_tokens(tb + 2 * BK, aPh, aPl, aV + sv * T64, idesc_o, false, leader);
        umma_commit_p(&v_empty[sv], leader);
        umma_commit_p(o_full, leader);
      }
    }
  } else {
    // ===================== softmax / output: one thread per query row =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t aPh = smem_u32(sPh), aPl = smem_u32(sPl), amask = smem_u32(smask);
    const float scale = rsqrtf((float)DK), scale2 = LOG2E * scale;
    load_mask_words(smask, p.maskbits + (int64_t)b * p.mask_words, p.mask_words, threadIdx.x - 64, 128);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float m = -INFINITY, l = 0.f, alpha_prev = 1.f;   // m: running maximum of the RAW scores (the scale is folded into the FFMA)
    float O[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) O[c] = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1, 17);
      tcgen05_fence_after();
      uint32_t r[BK];
      tmem_ld32(lane_addr + st * BK, r);
      tmem_ld32(lane_addr + st * BK + 32, r + 32);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      const uint32_t w0 = lds32(amask + 8 * j), w1 = lds32(amask + 8 * j + 4);
      if ((w0 & w1) != 0xffffffffu) {   // warp-uniform: only tiles that contain masked / out-of-range keys pay for the selects
#pragma unroll
        for (int c = 0; c < BK; ++c)
          if (!(((c < 32 ? w0 : w1) >> (c & 31)) & 1u)) r[c] = 0xff800000u;   // -inf
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < BK; ++c) mx[c & 3] = fmaxf(mx[c & 3], __uint_as_float(r[c]));
      const float m_new = fmaxf(fmaxf(m, fmaxf(mx[0], mx[1])), fmaxf(mx[2], mx[3]));
      const float nms = -((m_new == -INFINITY) ? 0.f : m_new) * scale2;
      const float alpha = ex2(fmaf(m, scale2, nms));
      float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < BK; ++c) {
        const float pv = ex2(fmaf(__uint_as_float(r[c]), scale2, nms));
        r[c] = __float_as_uint(pv);
        rs[c & 3] += pv;
      }
      l = fmaf(l, alpha, (rs[0] + rs[1]) + (rs[2] + rs[3]));
      m = m_new;
      if (j > 0) {   // the P V product of the previous tile (it has also released the P tiles)
        mbar_wait(o_full, (j - 1) & 1, 18);
        tcgen05_fence_after();
        {   // PV columns [0,32) = P V_hi, [32,64) = P V_lo
          uint32_t t[32];
          tmem_ld32(lane_addr + 2 * BK, t);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) O[c] = fmaf(O[c], alpha_prev, __uint_as_float(t[c]));
          tmem_ld32(lane_addr + 2 * BK + 32, t);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) O[c] += __uint_as_float(t[c]);
        }
        tcgen05_fence_before();
      }
      if constexpr (ATM) {
        // r[0..31] <- hi pairs, r[32..63] <- lo pairs (pair c = keys 2c, 2c+1), in place: pair c only reads r[2c], r[2c+1]
        uint32_t lo[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          uint32_t hh;
          split2(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1]), hh, lo[c]);
          r[c] = hh;
        }
        tmem_st32(lane_addr + 192, r);
        tmem_st32(lane_addr + 224, lo);
        tmem_st_wait();
        tcgen05_fence_before();
      } else {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * ch + e]);
          uint4 hh, ll;
          split8(v, hh, ll);
          store_chunk(aPh, aPl, row, ch, hh, ll);
        }
        fence_proxy_async_smem();
      }
      mbar_arrive(p_full);
      alpha_prev = alpha;
    }
    mbar_wait(o_full, (nt - 1) & 1, 19);
    tcgen05_fence_after();
    {
      uint32_t t[32];
      tmem_ld32(lane_addr + 2 * BK, t);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) O[c] = fmaf(O[c], alpha_prev, __uint_as_float(t[c]));
      tmem_ld32(lane_addr + 2 * BK + 32, t);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) O[c] += __uint_as_float(t[c]);
    }
    if (qi < p.S) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      float* o = p.ctx + ((int64_t)b * p.S + qi) * p.H + h * DK;
#pragma unroll
      for (int c = 0; c < DK; c += 4)
        *reinterpret_cast<float4*>(o + c) = make_float4(O[c] * inv, O[c + 1] * inv, O[c + 2] * inv, O[c + 3] * inv);
      p.lse[(int64_t)slab * p.S + qi] = m * scale + logf(l);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// =====================================================================================================================
// backward, rows = queries: dQ
// =====================================================================================================================
constexpr int DQ_THREADS = 192;
constexpr int DQ_NK = 3, DQ_NV = 2;   // K tiles live until the dQ product of their tile, V tiles only until dP
constexpr int DQ_SMEM = 2 * T128 + (DQ_NK + DQ_NV) * T64 + 2 * T128;   // Q, G | K ring | V ring | dS hi, lo

template <bool ATM>
__global__ void __launch_bounds__(DQ_THREADS, 2)
fa_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_g,
                 const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t smask[MAX_MASK_WORDS + 2];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sG = sQ + T128;
  uint8_t* sK = sG + T128;
  uint8_t* sV = sK + DQ_NK * T64;
  uint8_t* sDh = sV + DQ_NV * T64;
  uint8_t* sDl = sDh + T128;
  uint64_t* bars = (uint64_t*)(sDl + T128);
  uint64_t* qg_full = bars;
  uint64_t* k_full = bars + 1;             // [3]
  uint64_t* k_empty = k_full + DQ_NK;      // [3]
  uint64_t* v_full = k_empty + DQ_NK;      // [2]
  uint64_t* v_empty = v_full + DQ_NV;      // [2]
  uint64_t* sd_full = v_empty + DQ_NV;
  uint64_t* sd_empty = sd_full + 1;
  uint64_t* ds_full = sd_empty + 1;
  uint64_t* ds_empty = ds_full + 1;
  uint64_t* acc_full = ds_empty + 1;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int slab = b * p.heads + h;
  const int nt = (p.S + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(qg_full, 1);
    for (int i = 0; i < DQ_NK; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < DQ_NV; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    mbar_init(sd_full, 1); mbar_init(sd_empty, 4);
    mbar_init(ds_full, 128); mbar_init(ds_empty, 1);
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // columns: S [0,64) | dP [64,128) | dQ [128,192) | dS hi [192,224) | dS lo [224,256)
  pdl_entry();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qg_full, 2 * T128);
      tma_load_3d(&map_q, qg_full, sQ, 0, q0, slab);
      tma_load_3d(&map_g, qg_full, sG, 0, q0, slab);
      for (int j = 0; j < nt; ++j) {
        const int sk = j % DQ_NK, sv = j % DQ_NV;
        mbar_wait(&k_empty[sk], ((j / DQ_NK) & 1) ^ 1, 20);
        mbar_expect_tx(&k_full[sk], T64);
        tma_load_3d(&map_k, &k_full[sk], sK + sk * T64, 0, j * BK, slab);
        mbar_wait(&v_empty[sv], ((j / DQ_NV) & 1) ^ 1, 21);
        mbar_expect_tx(&v_full[sv], T64);
        tma_load_3d(&map_v, &v_full[sv], sV + sv * T64, 0, j * BK, slab);
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t leader = elect_leader();
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0), sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t aQ = sb, aG = sb + T128, aK = aG + T128, aV = aK + DQ_NK * T64, aDh = aV + DQ_NV * T64, aDl = aDh + T128;
      const uint32_t idesc_s = make_idesc(BK, true, true, 2), idesc_o = make_idesc(64, true, false, 2);
      auto issue_sd = [&](int j) {
        const int sk = j % DQ_NK, sv = j % DQ_NV;
        mbar_wait(&k_full[sk], (j / DQ_NK) & 1, 22);
        mbar_wait(&v_full[sv], (j / DQ_NV) & 1, 23);
        mbar_wait(sd_empty, (j & 1) ^ 1, 24);
        tcgen05_fence_after();
        mma_rowsplit(tb, aQ, aK + sk * T64, idesc_s, leader);        // S = Q K^T
        mma_rowsplit(tb + BK, aG, aV + sv * T64, idesc_s, leader);   // dP = dO V^T
        umma_commit_p(&v_empty[sv], leader);
        umma_commit_p(sd_full, leader);
      };
      mbar_wait(qg_full, 0, 25);
      issue_sd(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_sd(j + 1);
        const int sk = j % DQ_NK;
        mbar_wait(ds_full, j & 1, 26);
        tcgen05_fence_after();
        if constexpr (ATM) mma_tokens_ts(tb + 2 * BK, tb + 192, tb + 224, aK + sk * T64, idesc_o, j > 0, leader);   // dQ += dS K
        else mma_tokens(tb + 2 * BK, aDh, aDl, aK + sk * T64, idesc_o, j > 0, leader);
        umma_commit_p(&k_empty[sk], leader);
        umma_commit_p(ds_empty, leader);
      }
      umma_commit_p(acc_full, leader);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t aDh = smem_u32(sDh), aDl = smem_u32(sDl), amask = smem_u32(smask);
    const float scale = rsqrtf((float)DK), scale2 = scale * LOG2E;
    // rows past S (Sq is a multiple of 64, q0 + row may still exceed it): -inf -> every weight 0
    const float NL = qi < p.Sq ? __ldg(p.L2 + (int64_t)slab * p.Sq + qi) : -INFINITY;   // -lse * log2(e)
    const float Dn = qi < p.Sq ? -__ldg(p.Dl + (int64_t)slab * p.Sq + qi) : 0.f;        // -delta / sqrt(dk)
    load_mask_words(smask, p.maskbits + (int64_t)b * p.mask_words, p.mask_words, threadIdx.x - 64, 128);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int j = 0; j < nt; ++j) {
      mbar_wait(sd_full, j & 1, 27);
      tcgen05_fence_after();
      uint32_t hh[32], ll[32];   // packed bf16 pairs of dS: pair c = keys 2c, 2c+1
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t rs_[32], rp[32];
        tmem_ld32(lane_addr + half * 32, rs_);
        tmem_ld32(lane_addr + BK + half * 32, rp);
        tmem_ld_wait();
        if (half == 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sd_empty);
        }
        const uint32_t w = lds32(amask + 4 * (2 * j + half));
        const bool full = w == 0xffffffffu;   // warp-uniform
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = 8 * ch + e;
            float pw = ex2(fmaf(__uint_as_float(rs_[c]), scale2, NL));
            if (!full) pw = ((w >> c) & 1u) ? pw : 0.f;
            v[e] = pw * fmaf(__uint_as_float(rp[c]), scale, Dn);
          }
          split8u(v, hh + 4 * (half * 4 + ch), ll + 4 * (half * 4 + ch));
        }
      }
      mbar_wait(ds_empty, (j & 1) ^ 1, 28);   // the dQ product of the previous tile has released the dS tiles
      if constexpr (ATM) {
        tcgen05_fence_after();
        tmem_st32(lane_addr + 192, hh);
        tmem_st32(lane_addr + 224, ll);
        tmem_st_wait();
        tcgen05_fence_before();
      } else {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          store_chunk(aDh, aDl, row, ch, make_uint4(hh[4 * ch], hh[4 * ch + 1], hh[4 * ch + 2], hh[4 * ch + 3]),
                      make_uint4(ll[4 * ch], ll[4 * ch + 1], ll[4 * ch + 2], ll[4 * ch + 3]));
        fence_proxy_async_smem();
      }
      mbar_arrive(ds_full);
    }
    mbar_wait(acc_full, 0, 29);
    tcgen05_fence_after();
    uint32_t t0[32], t1[32];
    tmem_ld32(lane_addr + 2 * BK, t0);
    tmem_ld32(lane_addr + 2 * BK + 32, t1);
    tmem_ld_wait();
    if (qi < p.S) {
      float* o = p.d_qkv + ((int64_t)b * p.S + qi) * (3 * (int64_t)p.H) + h * DK;
#pragma unroll
      for (int c = 0; c < DK; c += 4)
        *reinterpret_cast<float4*>(o + c) =
            make_float4(__uint_as_float(t0[c]) + __uint_as_float(t1[c]), __uint_as_float(t0[c + 1]) + __uint_as_float(t1[c + 1]),
                        __uint_as_float(t0[c + 2]) + __uint_as_float(t1[c + 2]), __uint_as_float(t0[c + 3]) + __uint_as_float(t1[c + 3]));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// =====================================================================================================================
// backward, rows = keys: dK, dV
// =====================================================================================================================
constexpr int DKV_THREADS = 320;   // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 columns 0..31, warps 6..9 columns 32..63
constexpr int DKV_NS_SMEM = 3, DKV_NS_TMEM = 6;   // ring depth: probabilities through shared memory / through tensor memory
constexpr int DKV_STAGE = 2 * T64 + 512;    // bytes landing per stage: Q, G tiles (row-split) | L2[64], D[64]
constexpr int DKV_PITCH = 2 * T64 + 1024;   // keeps every tile of every stage 1024-byte aligned
constexpr int DKV_SMEM_SMEM = 2 * T128 + DKV_NS_SMEM * DKV_PITCH + 8 * T128;   // K, V | ring | 2 x (P^T hi, lo, dS^T hi, lo)
constexpr int DKV_SMEM_TMEM = 2 * T128 + DKV_NS_TMEM * DKV_PITCH;             // K, V | ring

template <bool ATM>
__global__ void __launch_bounds__(DKV_THREADS, 1)
fa_bwd_dkv_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                  const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_g, const Params p) {
  constexpr int DKV_NS = ATM ? DKV_NS_TMEM : DKV_NS_SMEM;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = sK + T128;
  uint8_t* sSt = sV + T128;                 // [DKV_NS] stages
  uint8_t* sPD = sSt + DKV_NS * DKV_PITCH;  // !ATM: [2] x (P^T hi | P^T lo | dS^T hi | dS^T lo): the softmax warps fill one
                                            // set while the dV / dK products of the previous tile still read the other
  uint64_t* bars = (uint64_t*)(sPD + (ATM ? 0 : 8 * T128));
  uint64_t* kv_full = bars;
  uint64_t* t_full = bars + 1;              // [DKV_NS]
  uint64_t* t_empty = t_full + DKV_NS;      // [DKV_NS]
  uint64_t* sd_full = t_empty + DKV_NS;     // [2]
  uint64_t* sd_empty = sd_full + 2;         // [2]
  uint64_t* pd_full = sd_empty + 2;         // [2]
  uint64_t* pd_empty = pd_full + 2;         // [2]
  uint64_t* acc_full = pd_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * BQ;
  const int slab = b * p.heads + h;
  const int nq = (p.S + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < DKV_NS; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sd_full[i], 1); mbar_init(&sd_empty[i], 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(&pd_full[i], 256); mbar_init(&pd_empty[i], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // columns: [S^T | dP^T] x 2 buffers = [0,256) | dK [256,320) | dV [320,384) | ATM: P^T hi, lo [384,448) | dS^T hi, lo [448,512)
  const uint32_t tmem_base = *tmem_slot;
  pdl_entry();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * T128);
      tma_load_3d(&map_k, kv_full, sK, 0, k0, slab);
      tma_load_3d(&map_v, kv_full, sV, 0, k0, slab);
      for (int j = 0; j < nq; ++j) {
        const int st = j % DKV_NS;
        uint8_t* s = sSt + st * DKV_PITCH;
        mbar_wait(&t_empty[st], ((j / DKV_NS) & 1) ^ 1, 30);
        mbar_expect_tx(&t_full[st], DKV_STAGE);
        tma_load_3d(&map_q, &t_full[st], s, 0, j * BK, slab);
        tma_load_3d(&map_g, &t_full[st], s + T64, 0, j * BK, slab);
        bulk_load_1d(s + 2 * T64, p.L2 + (int64_t)slab * p.Sq + j * BK, 256, &t_full[st]);
        bulk_load_1d(s + 2 * T64 + 256, p.Dl + (int64_t)slab * p.Sq + j * BK, 256, &t_full[st]);
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t leader = elect_leader();
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0), sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t aK = sb, aV = sb + T128, aSt = aV + T128, aPD = aSt + DKV_NS * DKV_PITCH;
      const uint32_t idesc_s = make_idesc(BK, true, true, 2), idesc_o = make_idesc(64, true, false, 2);
      auto issue_sd = [&](int j) {
        const int tb = j & 1, st = j % DKV_NS;
        const uint32_t s = aSt + st * DKV_PITCH;
        FA_STAMP(0, j);
        mbar_wait(&t_full[st], (j / DKV_NS) & 1, 31);
        FA_STAMP(1, j);
        mbar_wait(&sd_empty[tb], ((j >> 1) & 1) ^ 1, 32);
        tcgen05_fence_after();
        FA_STAMP(2, j);
        mma_rowsplit(tm + tb * 2 * BK, aK, s, idesc_s, leader);              // S^T = K Q^T
        mma_rowsplit(tm + tb * 2 * BK + BK, aV, s + T64, idesc_s, leader);   // dP^T = V dO^T
        umma_commit_p(&sd_full[tb], leader);
        FA_STAMP(3, j);
      };
      mbar_wait(kv_full, 0, 33);
      issue_sd(0);
      if (ATM && nq > 1) issue_sd(1);
      for (int j = 0; j < nq; ++j) {
        // ATM (deep ring): two tiles ahead -- S^T / dP^T of tile j+2 reuse the buffers of tile j, free once the softmax warps
        // have loaded it; otherwise one tile ahead
        if (ATM) { if (j + 2 < nq) issue_sd(j + 2); }
        else if (j + 1 < nq) issue_sd(j + 1);
        const int st = j % DKV_NS, pb = j & 1;
        const uint32_t s = aSt + st * DKV_PITCH;
        const uint32_t pd = aPD + pb * 4 * T128;
        if constexpr (ATM) {   // one probability set in tensor memory: barriers [0] only, one completion per tile
          mbar_wait(&pd_full[0], j & 1, 34);
          tcgen05_fence_after();
          FA_STAMP(4, j);
          mma_tokens_ts(tm + 4 * BK + 64, tm + 384, tm + 416, s + T64, idesc_o, j > 0, leader);   // dV += P^T dO
          mma_tokens_ts(tm + 4 * BK, tm + 448, tm + 480, s, idesc_o, j > 0, leader);              // dK += dS^T Q
          umma_commit_p(&t_empty[st], leader);
          umma_commit_p(&pd_empty[0], leader);
          FA_STAMP(5, j);
        } else {
          mbar_wait(&pd_full[pb], (j >> 1) & 1, 34);
          tcgen05_fence_after();
          mma_tokens(tm + 4 * BK + 64, pd, pd + T128, s + T64, idesc_o, j > 0, leader);            // dV += P^T dO
          mma_tokens(tm + 4 * BK, pd + 2 * T128, pd + 3 * T128, s, idesc_o, j > 0, leader);        // dK += dS^T Q
          umma_commit_p(&t_empty[st], leader);
          umma_commit_p(&pd_empty[pb], leader);
        }
      }
      umma_commit_p(acc_full, leader);
    }
  } else {
    const int q = warp & 3;
    const int wg = (warp - 2) >> 2;          // 0: columns 0..31 of every tile, 1: columns 32..63
    const int row = q * 32 + lane;
    const int ki = k0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float scale = rsqrtf((float)DK), scale2 = scale * LOG2E;
    const bool key_ok = ki < p.S && ((__ldg(p.maskbits + (int64_t)b * p.mask_words + (ki >> 5)) >> (ki & 31)) & 1u);
    const bool all_ok = __all_sync(0xffffffffu, key_ok);   // warp-uniform: no per-element selects for fully valid warps
    for (int j = 0; j < nq; ++j) {
      const int tb = j & 1, st = j % DKV_NS;
      const uint32_t sstat = smem_u32(sSt + st * DKV_PITCH + 2 * T64) + wg * 128;   // this warpgroup's 32 queries
      if (threadIdx.x == 64) FA_STAMP(8, j);
      mbar_wait(&t_full[st], (j / DKV_NS) & 1, 38);      // the stage's L2 / D rows (bulk copies) are read by these threads
      mbar_wait(&sd_full[tb], (j >> 1) & 1, 35);
      tcgen05_fence_after();
      if (threadIdx.x == 64) FA_STAMP(9, j);
      uint32_t rs_[32], rp[32];
      tmem_ld32(lane_addr + tb * 2 * BK + wg * 32, rs_);
      tmem_ld32(lane_addr + tb * 2 * BK + BK + wg * 32, rp);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sd_empty[tb]);
      if (threadIdx.x == 64) FA_STAMP(10, j);
      uint32_t ph[16], pl[16], dh[16], dl[16];   // packed bf16 pairs of this warpgroup's 32 queries
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float pv[8], dv[8];
#pragma unroll
        for (int e4 = 0; e4 < 2; ++e4) {
          const float4 L4 = lds128f(sstat + (2 * ch + e4) * 16);          // -lse * log2(e) of four queries
          const float4 D4 = lds128f(sstat + 256 + (2 * ch + e4) * 16);    // delta / sqrt(dk)
          const float Lx[4] = {L4.x, L4.y, L4.z, L4.w}, Dx[4] = {D4.x, D4.y, D4.z, D4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 8 * ch + 4 * e4 + e;
            float pw = ex2(fmaf(__uint_as_float(rs_[c]), scale2, Lx[e]));
            if (!all_ok) pw = key_ok ? pw : 0.f;
            pv[4 * e4 + e] = pw;
            dv[4 * e4 + e] = pw * fmaf(__uint_as_float(rp[c]), scale, -Dx[e]);
          }
        }
        split8u(pv, ph + 4 * ch, pl + 4 * ch);
        split8u(dv, dh + 4 * ch, dl + 4 * ch);
      }
      if constexpr (ATM) {
        if (threadIdx.x == 64) FA_STAMP(11, j);
        mbar_wait(&pd_empty[0], (j & 1) ^ 1, 36);   // the dV / dK products of the previous tile have read the probabilities
        tcgen05_fence_after();
        if (threadIdx.x == 64) FA_STAMP(12, j);
        tmem_st16(lane_addr + 384 + wg * 16, ph);
        tmem_st16(lane_addr + 416 + wg * 16, pl);
        tmem_st16(lane_addr + 448 + wg * 16, dh);
        tmem_st16(lane_addr + 480 + wg * 16, dl);
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(&pd_full[0]);
        if (threadIdx.x == 64) FA_STAMP(13, j);
      } else {
        const int pb = j & 1;
        const uint32_t aPh = smem_u32(sPD + pb * 4 * T128), aPl = aPh + T128, aDh = aPh + 2 * T128, aDl = aPh + 3 * T128;
        mbar_wait(&pd_empty[pb], ((j >> 1) & 1) ^ 1, 36);   // the dV / dK products of tile j - 2 have released this tile set
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          store_chunk(aPh, aPl, row, wg * 4 + ch, make_uint4(ph[4 * ch], ph[4 * ch + 1], ph[4 * ch + 2], ph[4 * ch + 3]),
                      make_uint4(pl[4 * ch], pl[4 * ch + 1], pl[4 * ch + 2], pl[4 * ch + 3]));
          store_chunk(aDh, aDl, row, wg * 4 + ch, make_uint4(dh[4 * ch], dh[4 * ch + 1], dh[4 * ch + 2], dh[4 * ch + 3]),
                      make_uint4(dl[4 * ch], dl[4 * ch + 1], dl[4 * ch + 2], dl[4 * ch + 3]));
        }
        fence_proxy_async_smem();
        mbar_arrive(&pd_full[pb]);
      }
    }
    mbar_wait(acc_full, 0, 37);
    tcgen05_fence_after();
    uint32_t t0[32], t1[32];
    tmem_ld32(lane_addr + 4 * BK + wg * 64, t0);       // wg 0: dK, wg 1: dV
    tmem_ld32(lane_addr + 4 * BK + wg * 64 + 32, t1);
    tmem_ld_wait();
    if (ki < p.S) {
      float* o = p.d_qkv + ((int64_t)b * p.S + ki) * (3 * (int64_t)p.H) + (1 + wg) * (int64_t)p.H + h * DK;
#pragma unroll
      for (int c = 0; c < DK; c += 4)
        *reinterpret_cast<float4*>(o + c) =
            make_float4(__uint_as_float(t0[c]) + __uint_as_float(t1[c]), __uint_as_float(t0[c + 1]) + __uint_as_float(t1[c + 1]),
                        __uint_as_float(t0[c + 2]) + __uint_as_float(t1[c + 2]), __uint_as_float(t0[c + 3]) + __uint_as_float(t1[c + 3]));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
// RS array [slabs][S][64] bf16 as a 3-D tensor: rows past S of a slab read as zeros (TMA out-of-bounds fill)
static int make_map_rs(CUtensorMap* map, const void* ptr, int S, uint64_t slabs, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  MVF_REQUIRE(fn != nullptr, MVF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {64, (cuuint64_t)S, slabs};
  cuuint64_t strides[2] = {128, (cuuint64_t)S * 128};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVF_REQUIRE(r == CUDA_SUCCESS, MVF_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d) S %d slabs %llu", (int)r, S,
              (unsigned long long)slabs);
  return MVF_OK;
}

struct WsLayout {
  size_t rs_bytes, stat_bytes, mask_bytes;
  size_t off_rs[4], off_L2, off_D, off_mask, total;
  int Sq, mask_words;
};
static WsLayout ws_layout(int B, int S, int heads) {
  WsLayout w;
  const size_t slabs = (size_t)B * heads;
  w.Sq = (int)round_up(S, 64);
  w.mask_words = cdiv(S, 32);
  auto al = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  w.rs_bytes = al(slabs * (size_t)S * 64 * 2);
  w.stat_bytes = al(slabs * (size_t)w.Sq * 4);
  w.mask_bytes = al((size_t)B * w.mask_words * 4);
  size_t off = 0;
  for (int i = 0; i < 4; ++i) { w.off_rs[i] = off; off += w.rs_bytes; }
  w.off_L2 = off; off += w.stat_bytes;
  w.off_D = off; off += w.stat_bytes;
  w.off_mask = off; off += w.mask_bytes;
  w.total = off;
  return w;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  MVF_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return MVF_OK;
}

}  // namespace fa

// MVF_ATTN_FA_TMEM: 1 (default) = probabilities through tensor memory ("TS" MMAs), 0 = through shared memory
static bool attn_fa_tmem() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MVF_ATTN_FA_TMEM");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}
// MVF_ATTN_FA: 0 = never, 1 (default) = S > 64 when the caller allows split operands, 2 = every S (tests)
static int attn_fa_mode() {
  const char* e = getenv("MVF_ATTN_FA");
  return e ? atoi(e) : 1;
}
size_t attention_fa_ws_bytes(int B, int S, int heads) { return fa::ws_layout(B, S, heads).total; }
bool attention_fa_ok(int dtype, int S, int dk, int H, bool allow_split, const void* ws, size_t ws_bytes, int B, int heads) {
  const int mode = attn_fa_mode();
  if (mode == 0 || !allow_split || ws == nullptr) return false;
  if (mode == 1 && S <= 64) return false;
  if (S > 32 * fa::MAX_MASK_WORDS) return false;
  return dtype == MVF_F32 && dk == fa::DK && H == heads * dk && H % 4 == 0 && tc_available() &&
         ws_bytes >= attention_fa_ws_bytes(B, S, heads) && ((uintptr_t)ws & 1023) == 0;
}

int attention_fa_fwd(int B, int S, int heads, const void* qkv, const float* keymask, void* ctx, float* lse, void* ws,
                     cudaStream_t st) {
  using namespace fa;
  const int H = heads * DK;
  const WsLayout w = ws_layout(B, S, heads);
  char* base = (char*)ws;
  bf16* QS = (bf16*)(base + w.off_rs[0]);
  bf16* KS = (bf16*)(base + w.off_rs[1]);
  bf16* VS = (bf16*)(base + w.off_rs[2]);
  uint32_t* mask = (uint32_t*)(base + w.off_mask);
  const float* x = (const float*)qkv;
  PrepArgs a;
  memset(&a, 0, sizeof(a));
  a.m[0] = PrepSrc{x, 3 * (int64_t)H, QS};
  a.m[1] = PrepSrc{x + H, 3 * (int64_t)H, KS};
  a.m[2] = PrepSrc{x + 2 * H, 3 * (int64_t)H, VS};
  a.n = 3;
  a.keymask = keymask;
  a.maskbits = mask;
  a.mask_words = w.mask_words;
  launch_k(fa_prep_kernel, dim3(cdiv(S, 32), heads, B), 256, 0, st, a, S, heads);
  MVF_CHECK_LAUNCH();
  const uint64_t slabs = (uint64_t)B * heads;
  CUtensorMap mq, mk, mv;
  MVF_TRY(make_map_rs(&mq, QS, S, slabs, BQ));
  MVF_TRY(make_map_rs(&mk, KS, S, slabs, BK));
  MVF_TRY(make_map_rs(&mv, VS, S, slabs, BK));
  static bool configured = false;
  constexpr int smem = FWD_SMEM + 1024 + 256;
  if (!configured) {
    MVF_TRY(set_smem(fa_fwd_kernel<true>, smem));
    MVF_TRY(set_smem(fa_fwd_kernel<false>, smem));
    configured = true;
  }
  Params p;
  memset(&p, 0, sizeof(p));
  p.S = S; p.Sq = w.Sq; p.heads = heads; p.H = H;
  p.mask_words = w.mask_words; p.maskbits = mask;
  p.ctx = (float*)ctx; p.lse = lse;
  if (attn_fa_tmem()) launch_k(fa_fwd_kernel<true>, dim3(cdiv(S, BQ), heads, B), FWD_THREADS, smem, st, mq, mk, mv, p);
  else launch_k(fa_fwd_kernel<false>, dim3(cdiv(S, BQ), heads, B), FWD_THREADS, smem, st, mq, mk, mv, p);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

int attention_fa_bwd(int B, int S, int heads, const void* qkv, const float* keymask, const void* ctx, const float* lse,
                     const void* d_ctx, void* d_qkv, void* ws, cudaStream_t st) {
  using namespace fa;
  const int H = heads * DK;
  const WsLayout w = ws_layout(B, S, heads);
  char* base = (char*)ws;
  bf16* QS = (bf16*)(base + w.off_rs[0]);
  bf16* KS = (bf16*)(base + w.off_rs[1]);
  bf16* VS = (bf16*)(base + w.off_rs[2]);
  bf16* GS = (bf16*)(base + w.off_rs[3]);
  float* L2 = (float*)(base + w.off_L2);
  float* Dl = (float*)(base + w.off_D);
  uint32_t* mask = (uint32_t*)(base + w.off_mask);
  const float* x = (const float*)qkv;
  PrepArgs a;
  memset(&a, 0, sizeof(a));
  a.m[0] = PrepSrc{x, 3 * (int64_t)H, QS};
  a.m[1] = PrepSrc{x + H, 3 * (int64_t)H, KS};
  a.m[2] = PrepSrc{x + 2 * H, 3 * (int64_t)H, VS};
  a.m[3] = PrepSrc{(const float*)d_ctx, (int64_t)H, GS};
  a.n = 4;
  a.keymask = keymask;
  a.maskbits = mask;
  a.mask_words = w.mask_words;
  launch_k(fa_prep_kernel, dim3(cdiv(S, 32), heads, B), 256, 0, st, a, S, heads);
  MVF_CHECK_LAUNCH();
  launch_k(fa_stats_kernel, dim3(w.Sq / 32, heads, B), 256, 0, st, (const float*)ctx, (const float*)d_ctx, lse, L2, Dl, S, w.Sq, heads, H);
  MVF_CHECK_LAUNCH();
  const uint64_t slabs = (uint64_t)B * heads;
  CUtensorMap mq128, mg128, mk64, mv64, mk128, mv128, mq64, mg64;
  MVF_TRY(make_map_rs(&mq128, QS, S, slabs, BQ));
  MVF_TRY(make_map_rs(&mg128, GS, S, slabs, BQ));
  MVF_TRY(make_map_rs(&mk64, KS, S, slabs, BK));
  MVF_TRY(make_map_rs(&mv64, VS, S, slabs, BK));
  MVF_TRY(make_map_rs(&mk128, KS, S, slabs, BQ));
  MVF_TRY(make_map_rs(&mv128, VS, S, slabs, BQ));
  MVF_TRY(make_map_rs(&mq64, QS, S, slabs, BK));
  MVF_TRY(make_map_rs(&mg64, GS, S, slabs, BK));
  static bool configured = false;
  constexpr int smem_dq = DQ_SMEM + 1024 + 256;
  constexpr int smem_dkv_t = DKV_SMEM_TMEM + 1024 + 256, smem_dkv_s = DKV_SMEM_SMEM + 1024 + 256;
  if (!configured) {
    MVF_TRY(set_smem(fa_bwd_dq_kernel<true>, smem_dq));
    MVF_TRY(set_smem(fa_bwd_dq_kernel<false>, smem_dq));
    MVF_TRY(set_smem(fa_bwd_dkv_kernel<true>, smem_dkv_t));
    MVF_TRY(set_smem(fa_bwd_dkv_kernel<false>, smem_dkv_s));
    configured = true;
  }
  Params p;
  memset(&p, 0, sizeof(p));
  p.S = S; p.Sq = w.Sq; p.heads = heads; p.H = H;
  p.mask_words = w.mask_words; p.maskbits = mask;
  p.L2 = L2; p.Dl = Dl;
  p.d_qkv = (float*)d_qkv;
  static int dbg_on = -1;
  static long long* dbg_buf = nullptr;
  if (dbg_on < 0) {
    const char* e = getenv("MVF_FA_DBG");
    dbg_on = (e && atoi(e) != 0) ? 1 : 0;
    if (dbg_on && cudaMalloc(&dbg_buf, 24 * 16 * sizeof(long long)) != cudaSuccess) dbg_on = 0;   // debug aid only
  }
  if (dbg_on) {
    cudaMemsetAsync(dbg_buf, 0, 24 * 16 * sizeof(long long), st);
    p.dbg = dbg_buf;
  }
  if (attn_fa_tmem()) launch_k(fa_bwd_dq_kernel<true>, dim3(cdiv(S, BQ), heads, B), DQ_THREADS, smem_dq, st, mq128, mg128, mk64, mv64, p);
  else launch_k(fa_bwd_dq_kernel<false>, dim3(cdiv(S, BQ), heads, B), DQ_THREADS, smem_dq, st, mq128, mg128, mk64, mv64, p);
  MVF_CHECK_LAUNCH();
  if (attn_fa_tmem()) launch_k(fa_bwd_dkv_kernel<true>, dim3(cdiv(S, BQ), heads, B), DKV_THREADS, smem_dkv_t, st, mk128, mv128, mq64, mg64, p);
  else launch_k(fa_bwd_dkv_kernel<false>, dim3(cdiv(S, BQ), heads, B), DKV_THREADS, smem_dkv_s, st, mk128, mv128, mq64, mg64, p);
  MVF_CHECK_LAUNCH();
  if (dbg_on) {
    long long h[24 * 16];
    if (cudaStreamSynchronize(st) == cudaSuccess && cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
      const long long t0 = h[0];
      fprintf(stderr, "fa_bwd_dkv dbg (S=%d): clocks relative to the first stamp; MMA warp: sd{enter,t_full,sd_empty,issued} dvdk{pd_full,issued} | "
              "softmax: enter sd_full ld_done compute_done pd_empty stored\n", S);
      for (int j = 0; j < 24 && j * 64 < S; ++j)
        fprintf(stderr, "  tile %2d: mma %7lld %7lld %7lld %7lld | %7lld %7lld || soft %7lld %7lld %7lld %7lld %7lld %7lld\n", j, h[j * 16 + 0] - t0,
                h[j * 16 + 1] - t0, h[j * 16 + 2] - t0, h[j * 16 + 3] - t0, h[j * 16 + 4] - t0, h[j * 16 + 5] - t0, h[j * 16 + 8] - t0,
                h[j * 16 + 9] - t0, h[j * 16 + 10] - t0, h[j * 16 + 11] - t0, h[j * 16 + 12] - t0, h[j * 16 + 13] - t0);
    }
  }
  return MVF_OK;
}

}  // namespace mvf

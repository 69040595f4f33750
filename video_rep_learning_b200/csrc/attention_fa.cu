// Temporal self-attention for long sequences (S > 64, d_k = 32) on the 5th-generation tensor cores: flash-attention style
// forward and backward with tcgen05.mma (accumulators in TMEM), operands staged by 3-D TMA tensor maps, online softmax in
// the exp2 domain.  models/utils.py:11-44 + 47-108 with the [B,1,1,S] key mask; S = E*T is 480 at BASELINE cfg4, 3840 at
// cfg5 and up to ~12 k in whole-video evaluation (evaluate.py:45-62), where the reference materialises [B,8,S,S] scores.
//
// Precision: the chain behind the pooling feeds SCL's 1/temperature, so every contraction uses bf16 hi/lo operand splits
// (x = hi + lo; hi*hi + hi*lo + lo*hi, ~16 mantissa bits, fp32 accumulation in TMEM) exactly like the forward GEMMs
// (gemm_tc.cu SPLIT3) and the S <= 64 kernels (attention_tc.cu).  A pre-pass writes the split operands once per layer:
//   "row-split"   RS_x [slab][S][64]    bf16: columns 0..31 = hi(x[s, :]), 32..63 = lo        (K-major A/B tiles, K = d_k)
//   "transposed"  TS_x [slab][64][Sp]   bf16: rows 0..31 = hi(x[:, d]), 32..63 = lo            (K-major B tiles, K = tokens)
// with slab = view * heads + head; one 128-byte row = one SWIZZLE_128B atom row, so the three products of a k-step are the
// same descriptors advanced by 0 / 64 bytes (row-split) or 0 / 32 rows (transposed).
//
//   forward   CTA = 128 queries x one slab; loop over 64-key tiles:  S = Q K^T (TMEM, double buffered) -> softmax warps
//             (one thread per row, tcgen05.ld) -> P hi|lo into shared memory -> PV = P V (TMEM) -> O kept in registers
//             (O = O * alpha + PV).  2 CTAs per SM.
//   dQ        CTA = 128 queries; per key tile S = Q K^T and dP = dO V^T (TMEM) -> dS = P (dP - delta) / sqrt(dk) hi|lo
//             into shared memory -> dQ += dS K accumulated in TMEM over the whole loop.  2 CTAs per SM.
//   dK, dV    CTA = 128 keys; per 64-query tile S^T = K Q^T, dP^T = V dO^T (TMEM, double buffered) -> P^T, dS^T hi|lo
//             into shared memory (two softmax warpgroups, 32 columns each) -> dV += P^T dO, dK += dS^T Q in TMEM.
// Every score is recomputed in the orientation its consumer needs, so no transposed accumulator travels through memory
// (the scheme of attention_tc.cu, now streaming).  d_k = 32 makes the kernels exp/ALU-bound, not tensor-bound
// (128 tensor FLOPs per exp): the roofline reported for them is MUFU ex2 throughput.
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
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 2^x on the MUFU (ex2.approx.ftz: 2^-inf = +0); exp2f() wraps the same instruction in range fix-ups the kernels do not need
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void split8(const float* v, uint4& h, uint4& l) {
  split2(v[0], v[1], h.x, l.x);
  split2(v[2], v[3], h.y, l.y);
  split2(v[4], v[5], h.z, l.z);
  split2(v[6], v[7], h.w, l.w);
}
__device__ __forceinline__ void store_chunk(uint8_t* tile_hi, uint8_t* tile_lo, int r, int chunk, const uint4& h, const uint4& l) {
  const int off = r * 128 + ((chunk ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(tile_hi + off) = h;
  *reinterpret_cast<uint4*>(tile_lo + off) = l;
}
constexpr int MAX_MASK_WORDS = 512;    // S <= 16384
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
// pre-pass: fp32 [B*S, ld] (head h at columns h*32 of each source) -> row-split / transposed bf16 hi|lo operands
// ---------------------------------------------------------------------------------------------------------------------
struct PrepSrc {
  const float* src;   // first row of view 0, first column of head 0
  int64_t ld;         // row stride (floats)
  bf16* rs;           // [slab][S][64] or null
  bf16* ts;           // [slab][64][Sp] or null
};
struct PrepArgs {
  PrepSrc m[4];
  int n;
  const float* keymask;   // [B, S] or null
  uint32_t* maskbits;     // [B][mask_words] or null: bit j of word w = key 32 w + j may be attended
  int mask_words;
};

__global__ void __launch_bounds__(256) fa_prep_kernel(const PrepArgs a, int S, int Sp, int heads) {
  pdl_entry();
  __shared__ uint16_t th[DK][40], tl[DK][40];   // [d][token] transposition tiles (hi, lo)
  const int b = blockIdx.z, h = blockIdx.y, s0 = blockIdx.x * 32, tid = threadIdx.x;
  const int64_t slab = (int64_t)b * heads + h;
  const int r = tid >> 3, c4 = tid & 7;
  const int s = s0 + r;
  if (a.maskbits != nullptr && h == 0 && tid < 32) {
    const int j = s0 + tid;
    const bool ok = j < S && (a.keymask == nullptr || a.keymask[(int64_t)b * S + j] != 0.f);
    const uint32_t w = __ballot_sync(0xffffffffu, ok);
    if (tid == 0) a.maskbits[(int64_t)b * a.mask_words + blockIdx.x] = w;
  }
  for (int i = 0; i < a.n; ++i) {
    const PrepSrc& m = a.m[i];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < S) v = *reinterpret_cast<const float4*>(m.src + ((int64_t)b * S + s) * m.ld + h * DK + 4 * c4);
    uint2 hi, lo;
    split2(v.x, v.y, hi.x, lo.x);
    split2(v.z, v.w, hi.y, lo.y);
    if (m.rs != nullptr && s < S) {
      bf16* row = m.rs + (slab * S + s) * 64;
      *reinterpret_cast<uint2*>(row + 4 * c4) = hi;
      *reinterpret_cast<uint2*>(row + 32 + 4 * c4) = lo;
    }
    if (m.ts != nullptr) {
      __syncthreads();   // previous matrix's tile fully consumed
      th[4 * c4 + 0][r] = (uint16_t)(hi.x & 0xffffu); th[4 * c4 + 1][r] = (uint16_t)(hi.x >> 16);
      th[4 * c4 + 2][r] = (uint16_t)(hi.y & 0xffffu); th[4 * c4 + 3][r] = (uint16_t)(hi.y >> 16);
      tl[4 * c4 + 0][r] = (uint16_t)(lo.x & 0xffffu); tl[4 * c4 + 1][r] = (uint16_t)(lo.x >> 16);
      tl[4 * c4 + 2][r] = (uint16_t)(lo.y & 0xffffu); tl[4 * c4 + 3][r] = (uint16_t)(lo.y >> 16);
      __syncthreads();
      // thread (d = r, token group c4): 4 tokens = 8 bytes of row d (hi) and row 32 + d (lo); tokens >= S are zeros
      const int t0 = s0 + 4 * c4;
      if (t0 < Sp) {
        uint2 oh, ol;
        oh.x = (uint32_t)th[r][4 * c4] | ((uint32_t)th[r][4 * c4 + 1] << 16);
        oh.y = (uint32_t)th[r][4 * c4 + 2] | ((uint32_t)th[r][4 * c4 + 3] << 16);
        ol.x = (uint32_t)tl[r][4 * c4] | ((uint32_t)tl[r][4 * c4 + 1] << 16);
        ol.y = (uint32_t)tl[r][4 * c4 + 2] | ((uint32_t)tl[r][4 * c4 + 3] << 16);
        bf16* base = m.ts + slab * 64 * (int64_t)Sp;
        *reinterpret_cast<uint2*>(base + (int64_t)r * Sp + t0) = oh;
        *reinterpret_cast<uint2*>(base + (int64_t)(32 + r) * Sp + t0) = ol;
      }
    }
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
// shared pieces of the three kernels
// ---------------------------------------------------------------------------------------------------------------------
// three products of one [128 x 32] x [N x 32]^T contraction on row-split operands (A row = hi32|lo32, B row = hi32|lo32)
__device__ __forceinline__ void mma_rowsplit(uint32_t tmem_d, uint32_t sa, uint32_t sb, uint32_t idesc) {
  const uint64_t da = make_smem_desc(sa, true, 0, 2), db = make_smem_desc(sb, true, 0, 2);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint64_t hi = (uint64_t)(2 * j), lo = (uint64_t)(4 + 2 * j);   // 16-byte units inside the 128-byte row
    umma<2>(tmem_d, da + lo, db + hi, idesc, j > 0 ? 1u : 0u);           // small terms first
    umma<2>(tmem_d, da + hi, db + lo, idesc, 1u);
    umma<2>(tmem_d, da + hi, db + hi, idesc, 1u);
  }
}
// D[128 x 32] (+)= A[128 x 64] * B[64 x 32] with A = hi|lo tiles in shared memory ([128 rows][64 k] bf16 each, 16 KB) and
// B = a transposed-split tile ([64 rows = hi d | lo d][64 k], 8 KB)
__device__ __forceinline__ void mma_tsplit(uint32_t tmem_d, uint32_t sa_hi, uint32_t sa_lo, uint32_t sb, uint32_t idesc, bool accumulate) {
  const uint64_t dah = make_smem_desc(sa_hi, true, 0, 2), dal = make_smem_desc(sa_lo, true, 0, 2);
  const uint64_t dbh = make_smem_desc(sb, true, 0, 2), dbl = make_smem_desc(sb + 32 * ROW, true, 0, 2);
#pragma unroll
  for (int k = 0; k < 4; ++k) {                                          // 4 x 16 tokens of K
    const uint64_t o = (uint64_t)(2 * k);
    umma<2>(tmem_d, dal + o, dbh + o, idesc, (accumulate || k > 0) ? 1u : 0u);
    umma<2>(tmem_d, dah + o, dbl + o, idesc, 1u);
    umma<2>(tmem_d, dah + o, dbh + o, idesc, 1u);
  }
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
};

// =====================================================================================================================
// forward
// =====================================================================================================================
constexpr int FWD_THREADS = 192;   // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 softmax (TMEM lane quadrant = warp % 4)
constexpr int FWD_SMEM = 16384 + 2 * 8192 + 2 * 8192 + 2 * 16384;

__global__ void __launch_bounds__(FWD_THREADS, 2)
fa_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
              const __grid_constant__ CUtensorMap map_vt, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t smask[MAX_MASK_WORDS + 2];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 16384;
  uint8_t* sV = sK + 2 * 8192;
  uint8_t* sPh = sV + 2 * 8192;
  uint8_t* sPl = sPh + 16384;
  uint64_t* bars = (uint64_t*)(sPl + 16384);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]
  uint64_t* s_full = bars + 9;     // [2]
  uint64_t* s_empty = bars + 11;   // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* o_full = bars + 14;
  uint32_t* tmem_slot = (uint32_t*)(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int slab = b * p.heads + h;
  const int nt = (p.S + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_vt) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
    }
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
  const uint32_t tmem_base = *tmem_slot;   // columns: S0 [0,64) | S1 [64,128) | PV [128,160)
  pdl_entry();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 16384);
      tma_load_3d(&map_q, q_full, sQ, 0, q0, slab);
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1, 10);
        mbar_expect_tx(&k_full[st], 8192);
        tma_load_3d(&map_k, &k_full[st], sK + st * 8192, 0, j * BK, slab);
        mbar_wait(&v_empty[st], ph ^ 1, 11);
        mbar_expect_tx(&v_full[st], 8192);
        tma_load_3d(&map_vt, &v_full[st], sV + st * 8192, j * BK, 0, slab);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(BK, true, true, 2), idesc_o = make_idesc(DK, true, true, 2);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[st], ph, 12);
        mbar_wait(&s_empty[st], ph ^ 1, 13);
        tcgen05_fence_after();
        mma_rowsplit(tmem_base + st * BK, smem_u32(sQ), smem_u32(sK + st * 8192), idesc_s);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
      };
      mbar_wait(q_full, 0, 14);
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_s(j + 1);
        const int st = j & 1;
        mbar_wait(p_full, j & 1, 15);
        mbar_wait(&v_full[st], (j >> 1) & 1, 16);
        tcgen05_fence_after();
        mma_tsplit(tmem_base + 2 * BK, smem_u32(sPh), smem_u32(sPl), smem_u32(sV + st * 8192), idesc_o, false);
        umma_commit(&v_empty[st]);
        umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax / output: one thread per query row =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
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
      const uint32_t w0 = smask[2 * j], w1 = smask[2 * j + 1];
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
        uint32_t t[DK];
        tmem_ld32(lane_addr + 2 * BK, t);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < DK; ++c) O[c] = fmaf(O[c], alpha_prev, __uint_as_float(t[c]));
        tcgen05_fence_before();
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * ch + e]);
        uint4 hh, ll;
        split8(v, hh, ll);
        store_chunk(sPh, sPl, row, ch, hh, ll);
      }
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      alpha_prev = alpha;
    }
    mbar_wait(o_full, (nt - 1) & 1, 19);
    tcgen05_fence_after();
    {
      uint32_t t[DK];
      tmem_ld32(lane_addr + 2 * BK, t);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < DK; ++c) O[c] = fmaf(O[c], alpha_prev, __uint_as_float(t[c]));
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
constexpr int DQ_SMEM = 2 * 16384 + 2 * 8192 + 2 * 8192 + 8192 + 2 * 16384;   // Q, G | K[2] | V[2] | KT | dS hi, lo

__global__ void __launch_bounds__(DQ_THREADS, 2)
fa_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_g,
                 const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                 const __grid_constant__ CUtensorMap map_kt, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t smask[MAX_MASK_WORDS + 2];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sG = sQ + 16384;
  uint8_t* sK = sG + 16384;
  uint8_t* sV = sK + 2 * 8192;
  uint8_t* sKT = sV + 2 * 8192;
  uint8_t* sDh = sKT + 8192;
  uint8_t* sDl = sDh + 16384;
  uint64_t* bars = (uint64_t*)(sDl + 16384);
  uint64_t* qg_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* kt_full = bars + 5;
  uint64_t* kt_empty = bars + 6;
  uint64_t* sd_full = bars + 7;
  uint64_t* sd_empty = bars + 8;
  uint64_t* ds_full = bars + 9;
  uint64_t* ds_empty = bars + 10;
  uint64_t* acc_full = bars + 11;
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int slab = b * p.heads + h;
  const int nt = (p.S + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kt) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(qg_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(kt_full, 1); mbar_init(kt_empty, 1);
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
  const uint32_t tmem_base = *tmem_slot;   // columns: S [0,64) | dP [64,128) | dQ [128,160)
  pdl_entry();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qg_full, 2 * 16384);
      tma_load_3d(&map_q, qg_full, sQ, 0, q0, slab);
      tma_load_3d(&map_g, qg_full, sG, 0, q0, slab);
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1, 20);
        mbar_expect_tx(&kv_full[st], 2 * 8192);
        tma_load_3d(&map_k, &kv_full[st], sK + st * 8192, 0, j * BK, slab);
        tma_load_3d(&map_v, &kv_full[st], sV + st * 8192, 0, j * BK, slab);
        mbar_wait(kt_empty, (j & 1) ^ 1, 21);
        mbar_expect_tx(kt_full, 8192);
        tma_load_3d(&map_kt, kt_full, sKT, j * BK, 0, slab);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(BK, true, true, 2), idesc_o = make_idesc(DK, true, true, 2);
      auto issue_sd = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1, 22);
        mbar_wait(sd_empty, (j & 1) ^ 1, 23);
        tcgen05_fence_after();
        mma_rowsplit(tmem_base, smem_u32(sQ), smem_u32(sK + st * 8192), idesc_s);        // S = Q K^T
        mma_rowsplit(tmem_base + BK, smem_u32(sG), smem_u32(sV + st * 8192), idesc_s);   // dP = dO V^T
        umma_commit(&kv_empty[st]);
        umma_commit(sd_full);
      };
      mbar_wait(qg_full, 0, 24);
      issue_sd(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_sd(j + 1);
        mbar_wait(ds_full, j & 1, 25);
        mbar_wait(kt_full, j & 1, 26);
        tcgen05_fence_after();
        mma_tsplit(tmem_base + 2 * BK, smem_u32(sDh), smem_u32(sDl), smem_u32(sKT), idesc_o, j > 0);   // dQ += dS K
        umma_commit(kt_empty);
        umma_commit(ds_empty);
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float scale = rsqrtf((float)DK), scale2 = scale * LOG2E;
    // rows past S (Sq is a multiple of 64, q0 + row may still exceed it): -inf -> every weight 0
    const float NL = qi < p.Sq ? __ldg(p.L2 + (int64_t)slab * p.Sq + qi) : -INFINITY;   // -lse * log2(e)
    const float Dn = qi < p.Sq ? -__ldg(p.Dl + (int64_t)slab * p.Sq + qi) : 0.f;        // -delta / sqrt(dk)
    load_mask_words(smask, p.maskbits + (int64_t)b * p.mask_words, p.mask_words, threadIdx.x - 64, 128);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int j = 0; j < nt; ++j) {
      mbar_wait(sd_full, j & 1, 27);
      tcgen05_fence_after();
      uint4 hh[8], ll[8];
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
        const uint32_t w = smask[2 * j + half];
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
          split8(v, hh[half * 4 + ch], ll[half * 4 + ch]);
        }
      }
      mbar_wait(ds_empty, (j & 1) ^ 1, 28);   // the dQ product of the previous tile has released the dS tiles
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) store_chunk(sDh, sDl, row, ch, hh[ch], ll[ch]);
      fence_proxy_async_smem();
      mbar_arrive(ds_full);
    }
    mbar_wait(acc_full, 0, 29);
    tcgen05_fence_after();
    uint32_t t[DK];
    tmem_ld32(lane_addr + 2 * BK, t);
    tmem_ld_wait();
    if (qi < p.S) {
      float* o = p.d_qkv + ((int64_t)b * p.S + qi) * (3 * (int64_t)p.H) + h * DK;
#pragma unroll
      for (int c = 0; c < DK; c += 4)
        *reinterpret_cast<float4*>(o + c) = make_float4(__uint_as_float(t[c]), __uint_as_float(t[c + 1]), __uint_as_float(t[c + 2]),
                                                        __uint_as_float(t[c + 3]));
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
constexpr int DKV_STAGE = 4 * 8192 + 512;   // bytes landing per stage: Q, G (row-split) | QT, GT (transposed) | L2[64], D[64]
constexpr int DKV_STAGE_PITCH = 34816;      // 34 KB: keeps every tile of every stage 1024-byte aligned
constexpr int DKV_SMEM = 2 * 16384 + 2 * DKV_STAGE_PITCH + 4 * 16384;   // K, V | stages | P^T hi, lo, dS^T hi, lo

__global__ void __launch_bounds__(DKV_THREADS, 1)
fa_bwd_dkv_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                  const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_g,
                  const __grid_constant__ CUtensorMap map_qt, const __grid_constant__ CUtensorMap map_gt, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = sK + 16384;
  uint8_t* sSt = sV + 16384;                 // [2] stages
  constexpr int STAGE_PITCH = DKV_STAGE_PITCH;
  uint8_t* sPh = sSt + 2 * STAGE_PITCH;
  uint8_t* sPl = sPh + 16384;
  uint8_t* sDh = sPl + 16384;
  uint8_t* sDl = sDh + 16384;
  uint64_t* bars = (uint64_t*)(sDl + 16384);
  uint64_t* kv_full = bars;
  uint64_t* t_full = bars + 1;     // [2]
  uint64_t* t_empty = bars + 3;    // [2]
  uint64_t* sd_full = bars + 5;    // [2]
  uint64_t* sd_empty = bars + 7;   // [2]
  uint64_t* pd_full = bars + 9;
  uint64_t* pd_empty = bars + 10;
  uint64_t* acc_full = bars + 11;
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * BQ;
  const int slab = b * p.heads + h;
  const int nq = (p.S + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qt) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_gt) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 1);
      mbar_init(&sd_full[i], 1); mbar_init(&sd_empty[i], 8);
    }
    mbar_init(pd_full, 256); mbar_init(pd_empty, 1);
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
  const uint32_t tmem_base = *tmem_slot;   // columns: [S^T | dP^T] x 2 stages = [0,256) | dK [256,288) | dV [288,320)
  pdl_entry();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * 16384);
      tma_load_3d(&map_k, kv_full, sK, 0, k0, slab);
      tma_load_3d(&map_v, kv_full, sV, 0, k0, slab);
      for (int j = 0; j < nq; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        uint8_t* s = sSt + st * STAGE_PITCH;
        mbar_wait(&t_empty[st], ph ^ 1, 30);
        mbar_expect_tx(&t_full[st], DKV_STAGE);
        tma_load_3d(&map_q, &t_full[st], s, 0, j * BK, slab);
        tma_load_3d(&map_g, &t_full[st], s + 8192, 0, j * BK, slab);
        tma_load_3d(&map_qt, &t_full[st], s + 2 * 8192, j * BK, 0, slab);
        tma_load_3d(&map_gt, &t_full[st], s + 3 * 8192, j * BK, 0, slab);
        bulk_load_1d(s + 4 * 8192, p.L2 + (int64_t)slab * p.Sq + j * BK, 256, &t_full[st]);
        bulk_load_1d(s + 4 * 8192 + 256, p.Dl + (int64_t)slab * p.Sq + j * BK, 256, &t_full[st]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(BK, true, true, 2), idesc_o = make_idesc(DK, true, true, 2);
      auto issue_sd = [&](int j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        uint8_t* s = sSt + st * STAGE_PITCH;
        mbar_wait(&t_full[st], ph, 31);
        mbar_wait(&sd_empty[st], ph ^ 1, 32);
        tcgen05_fence_after();
        mma_rowsplit(tmem_base + st * 2 * BK, smem_u32(sK), smem_u32(s), idesc_s);                 // S^T = K Q^T
        mma_rowsplit(tmem_base + st * 2 * BK + BK, smem_u32(sV), smem_u32(s + 8192), idesc_s);     // dP^T = V dO^T
        umma_commit(&sd_full[st]);
      };
      mbar_wait(kv_full, 0, 33);
      issue_sd(0);
      for (int j = 0; j < nq; ++j) {
        if (j + 1 < nq) issue_sd(j + 1);
        const int st = j & 1;
        uint8_t* s = sSt + st * STAGE_PITCH;
        mbar_wait(pd_full, j & 1, 34);
        tcgen05_fence_after();
        mma_tsplit(tmem_base + 4 * BK + DK, smem_u32(sPh), smem_u32(sPl), smem_u32(s + 3 * 8192), idesc_o, j > 0);   // dV += P^T dO
        mma_tsplit(tmem_base + 4 * BK, smem_u32(sDh), smem_u32(sDl), smem_u32(s + 2 * 8192), idesc_o, j > 0);        // dK += dS^T Q
        umma_commit(&t_empty[st]);
        umma_commit(pd_empty);
      }
      umma_commit(acc_full);
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
      const int st = j & 1;
      const uint8_t* s = sSt + st * STAGE_PITCH;
      mbar_wait(&t_full[st], (j >> 1) & 1, 38);      // the stage's L2 / D rows (bulk copies) are read by these threads
      mbar_wait(&sd_full[st], (j >> 1) & 1, 35);
      tcgen05_fence_after();
      uint32_t rs_[32], rp[32];
      tmem_ld32(lane_addr + st * 2 * BK + wg * 32, rs_);
      tmem_ld32(lane_addr + st * 2 * BK + BK + wg * 32, rp);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sd_empty[st]);
      const float4* Ls = reinterpret_cast<const float4*>(s + 4 * 8192) + wg * 8;          // -lse * log2(e) of the 32 queries
      const float4* Ds = reinterpret_cast<const float4*>(s + 4 * 8192 + 256) + wg * 8;    // delta / sqrt(dk)
      uint4 ph[4], pl[4], dh[4], dl[4];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float pv[8], dv[8];
#pragma unroll
        for (int e4 = 0; e4 < 2; ++e4) {
          const float4 L4 = Ls[2 * ch + e4], D4 = Ds[2 * ch + e4];
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
        split8(pv, ph[ch], pl[ch]);
        split8(dv, dh[ch], dl[ch]);
      }
      mbar_wait(pd_empty, (j & 1) ^ 1, 36);   // the dV / dK products of the previous tile have released the P^T / dS^T tiles
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        store_chunk(sPh, sPl, row, wg * 4 + ch, ph[ch], pl[ch]);
        store_chunk(sDh, sDl, row, wg * 4 + ch, dh[ch], dl[ch]);
      }
      fence_proxy_async_smem();
      mbar_arrive(pd_full);
    }
    mbar_wait(acc_full, 0, 37);
    tcgen05_fence_after();
    uint32_t t[DK];
    tmem_ld32(lane_addr + 4 * BK + wg * DK, t);       // wg 0: dK, wg 1: dV
    tmem_ld_wait();
    if (ki < p.S) {
      float* o = p.d_qkv + ((int64_t)b * p.S + ki) * (3 * (int64_t)p.H) + (1 + wg) * (int64_t)p.H + h * DK;
#pragma unroll
      for (int c = 0; c < DK; c += 4)
        *reinterpret_cast<float4*>(o + c) = make_float4(__uint_as_float(t[c]), __uint_as_float(t[c + 1]), __uint_as_float(t[c + 2]),
                                                        __uint_as_float(t[c + 3]));
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
static int make_map3(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                     uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = get_encode_fn();
  MVF_REQUIRE(fn != nullptr, MVF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVF_REQUIRE(r == CUDA_SUCCESS, MVF_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d) dims %llu x %llu x %llu", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2);
  return MVF_OK;
}

struct WsLayout {
  size_t rs_bytes, ts_bytes, stat_bytes, mask_bytes;
  size_t off_rs[4], off_ts[3], off_L2, off_D, off_mask, total;
  int Sp, Sq, mask_words;
};
static WsLayout ws_layout(int B, int S, int heads) {
  WsLayout w;
  const size_t slabs = (size_t)B * heads;
  w.Sp = (int)round_up(S, 8);
  w.Sq = (int)round_up(S, 64);
  w.mask_words = cdiv(S, 32);
  auto al = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  w.rs_bytes = al(slabs * (size_t)S * 64 * 2);
  w.ts_bytes = al(slabs * 64 * (size_t)w.Sp * 2);
  w.stat_bytes = al(slabs * (size_t)w.Sq * 4);
  w.mask_bytes = al((size_t)B * w.mask_words * 4);
  size_t off = 0;
  for (int i = 0; i < 4; ++i) { w.off_rs[i] = off; off += w.rs_bytes; }
  for (int i = 0; i < 3; ++i) { w.off_ts[i] = off; off += w.ts_bytes; }
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
  bf16* VT = (bf16*)(base + w.off_ts[0]);
  uint32_t* mask = (uint32_t*)(base + w.off_mask);
  const float* x = (const float*)qkv;
  PrepArgs a;
  memset(&a, 0, sizeof(a));
  a.m[0] = PrepSrc{x, 3 * (int64_t)H, QS, nullptr};
  a.m[1] = PrepSrc{x + H, 3 * (int64_t)H, KS, nullptr};
  a.m[2] = PrepSrc{x + 2 * H, 3 * (int64_t)H, nullptr, VT};
  a.n = 3;
  a.keymask = keymask;
  a.maskbits = mask;
  a.mask_words = w.mask_words;
  launch_k(fa_prep_kernel, dim3(cdiv(S, 32), heads, B), 256, 0, st, a, S, w.Sp, heads);
  MVF_CHECK_LAUNCH();
  const uint64_t slabs = (uint64_t)B * heads;
  CUtensorMap mq, mk, mvt;
  MVF_TRY(make_map3(&mq, QS, 64, S, slabs, 128, (uint64_t)S * 128, 64, BQ));
  MVF_TRY(make_map3(&mk, KS, 64, S, slabs, 128, (uint64_t)S * 128, 64, BK));
  MVF_TRY(make_map3(&mvt, VT, w.Sp, 64, slabs, (uint64_t)w.Sp * 2, (uint64_t)w.Sp * 128, 64, 64));
  static bool configured = false;
  constexpr int smem = FWD_SMEM + 1024 + 256;
  if (!configured) {
    MVF_TRY(set_smem(fa_fwd_kernel, smem));
    configured = true;
  }
  Params p;
  memset(&p, 0, sizeof(p));
  p.S = S; p.Sq = w.Sq; p.heads = heads; p.H = H;
  p.mask_words = w.mask_words; p.maskbits = mask;
  p.ctx = (float*)ctx; p.lse = lse;
  launch_k(fa_fwd_kernel, dim3(cdiv(S, BQ), heads, B), FWD_THREADS, smem, st, mq, mk, mvt, p);
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
  bf16* QT = (bf16*)(base + w.off_ts[0]);
  bf16* KT = (bf16*)(base + w.off_ts[1]);
  bf16* GT = (bf16*)(base + w.off_ts[2]);
  float* L2 = (float*)(base + w.off_L2);
  float* Dl = (float*)(base + w.off_D);
  uint32_t* mask = (uint32_t*)(base + w.off_mask);
  const float* x = (const float*)qkv;
  PrepArgs a;
  memset(&a, 0, sizeof(a));
  a.m[0] = PrepSrc{x, 3 * (int64_t)H, QS, QT};
  a.m[1] = PrepSrc{x + H, 3 * (int64_t)H, KS, KT};
  a.m[2] = PrepSrc{x + 2 * H, 3 * (int64_t)H, VS, nullptr};
  a.m[3] = PrepSrc{(const float*)d_ctx, (int64_t)H, GS, GT};
  a.n = 4;
  a.keymask = keymask;
  a.maskbits = mask;
  a.mask_words = w.mask_words;
  launch_k(fa_prep_kernel, dim3(cdiv(S, 32), heads, B), 256, 0, st, a, S, w.Sp, heads);
  MVF_CHECK_LAUNCH();
  launch_k(fa_stats_kernel, dim3(w.Sq / 32, heads, B), 256, 0, st, (const float*)ctx, (const float*)d_ctx, lse, L2, Dl, S, w.Sq, heads, H);
  MVF_CHECK_LAUNCH();
  const uint64_t slabs = (uint64_t)B * heads;
  const uint64_t rs2 = (uint64_t)S * 128, ts1 = (uint64_t)w.Sp * 2, ts2 = (uint64_t)w.Sp * 128;
  CUtensorMap mq128, mg128, mk64, mv64, mkt, mk128, mv128, mq64, mg64, mqt, mgt;
  MVF_TRY(make_map3(&mq128, QS, 64, S, slabs, 128, rs2, 64, BQ));
  MVF_TRY(make_map3(&mg128, GS, 64, S, slabs, 128, rs2, 64, BQ));
  MVF_TRY(make_map3(&mk64, KS, 64, S, slabs, 128, rs2, 64, BK));
  MVF_TRY(make_map3(&mv64, VS, 64, S, slabs, 128, rs2, 64, BK));
  MVF_TRY(make_map3(&mkt, KT, w.Sp, 64, slabs, ts1, ts2, 64, 64));
  MVF_TRY(make_map3(&mk128, KS, 64, S, slabs, 128, rs2, 64, BQ));
  MVF_TRY(make_map3(&mv128, VS, 64, S, slabs, 128, rs2, 64, BQ));
  MVF_TRY(make_map3(&mq64, QS, 64, S, slabs, 128, rs2, 64, BK));
  MVF_TRY(make_map3(&mg64, GS, 64, S, slabs, 128, rs2, 64, BK));
  MVF_TRY(make_map3(&mqt, QT, w.Sp, 64, slabs, ts1, ts2, 64, 64));
  MVF_TRY(make_map3(&mgt, GT, w.Sp, 64, slabs, ts1, ts2, 64, 64));
  static bool configured = false;
  constexpr int smem_dq = DQ_SMEM + 1024 + 256;
  constexpr int smem_dkv = DKV_SMEM + 1024 + 256;
  if (!configured) {
    MVF_TRY(set_smem(fa_bwd_dq_kernel, smem_dq));
    MVF_TRY(set_smem(fa_bwd_dkv_kernel, smem_dkv));
    configured = true;
  }
  Params p;
  memset(&p, 0, sizeof(p));
  p.S = S; p.Sq = w.Sq; p.heads = heads; p.H = H;
  p.mask_words = w.mask_words; p.maskbits = mask;
  p.L2 = L2; p.Dl = Dl;
  p.d_qkv = (float*)d_qkv;
  launch_k(fa_bwd_dq_kernel, dim3(cdiv(S, BQ), heads, B), DQ_THREADS, smem_dq, st, mq128, mg128, mk64, mv64, mkt, p);
  MVF_CHECK_LAUNCH();
  launch_k(fa_bwd_dkv_kernel, dim3(cdiv(S, BQ), heads, B), DKV_THREADS, smem_dkv, st, mk128, mv128, mq64, mg64, mqt, mgt, p);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace mvf

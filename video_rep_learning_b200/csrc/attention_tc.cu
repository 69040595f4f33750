// Temporal self-attention core for short sequences (S <= 64, d_k = 32) on the warp-level tensor cores.
//
// models/utils.py:11-44 + 47-108 at the named shapes: S = E*T = 60 tokens per view, 8 heads of 32 channels, key mask
// [B, S].  The CUDA-core "quad" kernels of attention.cu spend ~2400 instructions per thread on a 60 x 60 x 32 problem
// (25 us forward, 53 us backward per layer at B = 64: 11 % of the step); here one CTA of four warps owns one (view, head),
// every warp 16 rows, and all contractions are mma.sync.m16n8k16 on bf16 hi/lo splits of the fp32 operands
// (x = hi + lo, products hi*hi + hi*lo + lo*hi: ~16 mantissa bits, fp32 accumulation):
//
//   forward    S = Q K^T (log2 domain) -> masked softmax in registers -> O = P V          (P reused from the accumulators)
//   backward   phase A, warp = 16 queries:  S, dP = dO V^T, dS = P (dP - delta) / sqrt(dk), dQ = dS K
//              phase B, warp = 16 keys:     S^T = K Q^T, dP^T = V dO^T, dV = P^T dO, dK = dS^T Q
//              (S is recomputed in both orientations so that no transposed accumulator has to travel through memory)
//
// Operands are converted once per CTA into shared memory ([64 rows][32 ch] bf16, 80-byte rows: ldmatrix conflict-free);
// rows past S are zero.  Used by the fused head on the tensor-core backend (bf16 tokens); the exact-fp32 path keeps the
// CUDA-core kernels.
#include <math.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace mvf {
namespace atc {

constexpr int SP = 64;      // padded sequence length
constexpr int DK = 32;
constexpr int PITCH = 80;   // bytes per shared-memory row
constexpr int MAT = SP * PITCH;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// hi*hi + hi*lo + lo*hi
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
  mma(d, al, bh0, bh1);
  mma(d, ah, bl0, bl1);
  mma(d, ah, bh0, bh1);
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// [S rows][32 ch] fp32 (row stride ld) -> bf16 hi / lo matrices in shared memory; rows >= S are zero
__device__ __forceinline__ void load_split(const float* __restrict__ src, int64_t ld, int S, uint8_t* hi, uint8_t* lo) {
  for (int idx = threadIdx.x; idx < SP * 8; idx += blockDim.x) {
    const int row = idx >> 3, c4 = idx & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < S) v = *reinterpret_cast<const float4*>(src + (int64_t)row * ld + 4 * c4);
    uint2 h, l;
    split2(v.x, v.y, h.x, l.x);
    split2(v.z, v.w, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + row * PITCH + c4 * 8) = h;
    *reinterpret_cast<uint2*>(lo + row * PITCH + c4 * 8) = l;
  }
}

// A fragments (hi, lo) of rows [row0, row0+16) of a shared-memory matrix, both k-steps
__device__ __forceinline__ void load_a(const uint8_t* hi, const uint8_t* lo, int row0, int lane, uint32_t (&ah)[2][4],
                                       uint32_t (&al)[2][4]) {
  const uint32_t off = (uint32_t)((row0 + (lane & 15)) * PITCH + (lane >> 4) * 16);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    ldsm_x4(smem_u32(hi) + off + ks * 32, ah[ks]);
    ldsm_x4(smem_u32(lo) + off + ks * 32, al[ks]);
  }
}

// acc[nt] (16 x 8 tile nt of a 16 x 64 product) += A[16 x 32] * M[64 x 32]^T : the B operand is read row-wise from M
__device__ __forceinline__ void mm_nt(float (&acc)[8][4], const uint32_t (&ah)[2][4], const uint32_t (&al)[2][4], const uint8_t* mh,
                                      const uint8_t* ml, int lane) {
  const uint32_t off = (uint32_t)((lane & 7) * PITCH + (lane >> 3) * 16);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t bh[4], bl[4];   // {b0, b1} of k-step 0, {b0, b1} of k-step 1
    ldsm_x4(smem_u32(mh) + off + nt * 8 * PITCH, bh);
    ldsm_x4(smem_u32(ml) + off + nt * 8 * PITCH, bl);
    mma3(acc[nt], ah[0], al[0], bh[0], bh[1], bl[0], bl[1]);
    mma3(acc[nt], ah[1], al[1], bh[2], bh[3], bl[2], bl[3]);
  }
}

// out[nt] (16 x 8 tile nt of a 16 x 32 product) += P[16 x 64] * M[64 x 32] : P comes from accumulators (C layout == A layout
// of two adjacent 8-column tiles), the B operand is read column-wise from M (ldmatrix.trans)
__device__ __forceinline__ void mm_pv(float (&out)[4][4], const float (&p)[8][4], const uint8_t* mh, const uint8_t* ml, int lane) {
  const uint32_t off = (uint32_t)((lane & 15) * PITCH + (lane >> 4) * 16);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t ah[4], al[4];
    split2(p[2 * kk][0], p[2 * kk][1], ah[0], al[0]);
    split2(p[2 * kk][2], p[2 * kk][3], ah[1], al[1]);
    split2(p[2 * kk + 1][0], p[2 * kk + 1][1], ah[2], al[2]);
    split2(p[2 * kk + 1][2], p[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t bh[4], bl[4];   // {b0, b1} of channel tile 2*half, {b0, b1} of channel tile 2*half + 1
      ldsm_x4_trans(smem_u32(mh) + off + kk * 16 * PITCH + half * 32, bh);
      ldsm_x4_trans(smem_u32(ml) + off + kk * 16 * PITCH + half * 32, bl);
      mma3(out[2 * half], ah, al, bh[0], bh[1], bl[0], bl[1]);
      mma3(out[2 * half + 1], ah, al, bh[2], bh[3], bl[2], bl[3]);
    }
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_add(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// =====================================================================================================================
// forward: qkv [B*S, 3H] (Q | K | V, head h at columns h*32) -> ctx [B*S, H], lse [B, heads, S] (natural log)
// =====================================================================================================================
__global__ void __launch_bounds__(128)
attn_tc_fwd_kernel(int S, int H, const float* __restrict__ qkv, const float* __restrict__ keymask, float* __restrict__ ctx,
                   float* __restrict__ lse) {
  pdl_entry();
  __shared__ __align__(16) uint8_t sm[6 * MAT];   // Qh Ql Kh Kl Vh Vl
  __shared__ float Ms[SP];
  const int h = blockIdx.x, b = blockIdx.y, heads = gridDim.x, tid = threadIdx.x;
  const int64_t ld = 3 * (int64_t)H;
  const float* base = qkv + (int64_t)b * S * ld + h * DK;
  load_split(base, ld, S, sm, sm + MAT);
  load_split(base + H, ld, S, sm + 2 * MAT, sm + 3 * MAT);
  load_split(base + 2 * H, ld, S, sm + 4 * MAT, sm + 5 * MAT);
  for (int j = tid; j < SP; j += blockDim.x) Ms[j] = (j < S && (keymask == nullptr || keymask[(int64_t)b * S + j] != 0.f)) ? 1.f : 0.f;
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int row0 = warp * 16;
  if (row0 >= S) return;

  uint32_t qh[2][4], ql[2][4];
  load_a(sm, sm + MAT, row0, lane, qh, ql);
  float sc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
  mm_nt(sc, qh, ql, sm + 2 * MAT, sm + 3 * MAT, lane);

  const float scale2 = LOG2E / sqrtf((float)DK);
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const bool k0 = Ms[nt * 8 + 2 * t] != 0.f, k1 = Ms[nt * 8 + 2 * t + 1] != 0.f;
    sc[nt][0] = k0 ? sc[nt][0] * scale2 : -INFINITY;
    sc[nt][1] = k1 ? sc[nt][1] * scale2 : -INFINITY;
    sc[nt][2] = k0 ? sc[nt][2] * scale2 : -INFINITY;
    sc[nt][3] = k1 ? sc[nt][3] * scale2 : -INFINITY;
    m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1]));
    m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3]));
  }
  m0 = quad_max(m0);
  m1 = quad_max(m1);
  const float mm0 = m0 == -INFINITY ? 0.f : m0, mm1 = m1 == -INFINITY ? 0.f : m1;   // every key masked: all weights 0
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    sc[nt][0] = exp2f(sc[nt][0] - mm0); sc[nt][1] = exp2f(sc[nt][1] - mm0);
    sc[nt][2] = exp2f(sc[nt][2] - mm1); sc[nt][3] = exp2f(sc[nt][3] - mm1);
    l0 += sc[nt][0] + sc[nt][1];
    l1 += sc[nt][2] + sc[nt][3];
  }
  l0 = quad_add(l0);
  l1 = quad_add(l1);

  float o[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  mm_pv(o, sc, sm + 4 * MAT, sm + 5 * MAT, lane);

  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const int r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    if (r0 < S) *reinterpret_cast<float2*>(ctx + ((int64_t)b * S + r0) * H + h * DK + nt * 8 + 2 * t) = make_float2(o[nt][0] * i0, o[nt][1] * i0);
    if (r1 < S) *reinterpret_cast<float2*>(ctx + ((int64_t)b * S + r1) * H + h * DK + nt * 8 + 2 * t) = make_float2(o[nt][2] * i1, o[nt][3] * i1);
  }
  if (t == 0) {
    if (r0 < S) lse[((int64_t)b * heads + h) * S + r0] = m0 * 0.6931471805599453f + logf(l0);
    if (r1 < S) lse[((int64_t)b * heads + h) * S + r1] = m1 * 0.6931471805599453f + logf(l1);
  }
}

// =====================================================================================================================
// backward: d_ctx [B*S, H] -> d_qkv [B*S, 3H]
// =====================================================================================================================
// Eight warps: warps 0-3 take the query orientation (dQ), warps 4-7 the key orientation (dK, dV) of the same 16-row blocks --
// the two halves only share the staged operands, so the kernel's latency is one orientation, not the sum of both.
__global__ void __launch_bounds__(256)
attn_tc_bwd_kernel(int S, int H, const float* __restrict__ qkv, const float* __restrict__ keymask,
                   const float* __restrict__ ctx, const float* __restrict__ lse, const float* __restrict__ d_ctx,
                   float* __restrict__ d_qkv) {
  pdl_entry();
  __shared__ __align__(16) uint8_t sm[8 * MAT];   // Qh Ql Kh Kl Vh Vl Gh Gl   (G = dO)
  __shared__ float Ms[SP], Ls[SP], Ds[SP];
  const int h = blockIdx.x, b = blockIdx.y, heads = gridDim.x, tid = threadIdx.x;
  const int64_t ld = 3 * (int64_t)H;
  const float* base = qkv + (int64_t)b * S * ld + h * DK;
  const float* gbase = d_ctx + (int64_t)b * S * H + h * DK;
  const float* obase = ctx + (int64_t)b * S * H + h * DK;
  uint8_t *Qh = sm, *Ql = sm + MAT, *Kh = sm + 2 * MAT, *Kl = sm + 3 * MAT, *Vh = sm + 4 * MAT, *Vl = sm + 5 * MAT,
          *Gh = sm + 6 * MAT, *Gl = sm + 7 * MAT;
  {
    // all eight 16-byte loads of a thread are issued before the first conversion (one memory round trip for the four operands)
    static_assert(SP * 8 == 512, "two float4 per thread and operand at 256 threads");
    const float* srcs[4] = {base, base + H, base + 2 * H, gbase};
    const int64_t lds[4] = {ld, ld, ld, (int64_t)H};
    uint8_t* his[4] = {Qh, Kh, Vh, Gh};
    uint8_t* los[4] = {Ql, Kl, Vl, Gl};
    float4 v[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int idx = tid + u * 256, row = idx >> 3, c4 = idx & 7;
        v[a][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < S) v[a][u] = *reinterpret_cast<const float4*>(srcs[a] + (int64_t)row * lds[a] + 4 * c4);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int idx = tid + u * 256, row = idx >> 3, c4 = idx & 7;
        uint2 h, l;
        split2(v[a][u].x, v[a][u].y, h.x, l.x);
        split2(v[a][u].z, v[a][u].w, h.y, l.y);
        *reinterpret_cast<uint2*>(his[a] + row * PITCH + c4 * 8) = h;
        *reinterpret_cast<uint2*>(los[a] + row * PITCH + c4 * 8) = l;
      }
    }
  }
  // delta_i = <dO_i, O_i>: eight consecutive lanes own one row
  for (int idx = tid; idx < SP * 8; idx += blockDim.x) {
    const int row = idx >> 3, c4 = idx & 7;
    float dl = 0.f;
    if (row < S) {
      const float4 gv = *reinterpret_cast<const float4*>(gbase + (int64_t)row * H + 4 * c4);
      const float4 ov = *reinterpret_cast<const float4*>(obase + (int64_t)row * H + 4 * c4);
      dl = gv.x * ov.x + gv.y * ov.y + gv.z * ov.z + gv.w * ov.w;
    }
    dl += __shfl_xor_sync(0xffffffffu, dl, 1);
    dl += __shfl_xor_sync(0xffffffffu, dl, 2);
    dl += __shfl_xor_sync(0xffffffffu, dl, 4);
    if (c4 == 0) Ds[row] = dl;
  }
  for (int j = tid; j < SP; j += blockDim.x) {
    Ms[j] = (j < S && (keymask == nullptr || keymask[(int64_t)b * S + j] != 0.f)) ? 1.f : 0.f;
    // rows past S (and rows whose every key was masked: lse = -inf) get +inf so that 2^(s - L) = 0
    const float lv = j < S ? lse[((int64_t)b * heads + h) * S + j] : INFINITY;
    Ls[j] = (lv == -INFINITY) ? INFINITY : lv * LOG2E;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int row0 = (warp & 3) * 16;
  const bool keys_orientation = warp >= 4;
  if (row0 >= S) return;
  const float scale = 1.0f / sqrtf((float)DK);
  const float scale2 = scale * LOG2E;
  const int r0 = row0 + g, r1 = row0 + g + 8;

  // ---- phase A: rows = queries.  dQ_i = sum_j dS_ij K_j ----
  if (!keys_orientation) {
    uint32_t ah[2][4], al[2][4];
    float sc[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
    }
    load_a(Qh, Ql, row0, lane, ah, al);
    mm_nt(sc, ah, al, Kh, Kl, lane);        // S = Q K^T
    load_a(Gh, Gl, row0, lane, ah, al);
    mm_nt(dp, ah, al, Vh, Vl, lane);        // dP = dO V^T
    const float L0 = Ls[r0], L1 = Ls[r1], D0 = Ds[r0], D1 = Ds[r1];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const bool k0 = Ms[nt * 8 + 2 * t] != 0.f, k1 = Ms[nt * 8 + 2 * t + 1] != 0.f;
      const float p0 = k0 ? exp2f(sc[nt][0] * scale2 - L0) : 0.f, p1 = k1 ? exp2f(sc[nt][1] * scale2 - L0) : 0.f;
      const float p2 = k0 ? exp2f(sc[nt][2] * scale2 - L1) : 0.f, p3 = k1 ? exp2f(sc[nt][3] * scale2 - L1) : 0.f;
      sc[nt][0] = p0 * (dp[nt][0] - D0) * scale; sc[nt][1] = p1 * (dp[nt][1] - D0) * scale;
      sc[nt][2] = p2 * (dp[nt][2] - D1) * scale; sc[nt][3] = p3 * (dp[nt][3] - D1) * scale;
    }
    float dq[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
    mm_pv(dq, sc, Kh, Kl, lane);            // dQ = dS K
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (r0 < S) *reinterpret_cast<float2*>(d_qkv + ((int64_t)b * S + r0) * ld + h * DK + nt * 8 + 2 * t) = make_float2(dq[nt][0], dq[nt][1]);
      if (r1 < S) *reinterpret_cast<float2*>(d_qkv + ((int64_t)b * S + r1) * ld + h * DK + nt * 8 + 2 * t) = make_float2(dq[nt][2], dq[nt][3]);
    }
  }
  // ---- phase B: rows = keys.  dV_j = sum_i P_ij dO_i, dK_j = sum_i dS_ij Q_i ----
  else {
    uint32_t ah[2][4], al[2][4];
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
      dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
    }
    load_a(Kh, Kl, row0, lane, ah, al);
    mm_nt(st, ah, al, Qh, Ql, lane);        // S^T = K Q^T
    load_a(Vh, Vl, row0, lane, ah, al);
    mm_nt(dpt, ah, al, Gh, Gl, lane);       // dP^T = V dO^T
    const bool ka0 = Ms[r0] != 0.f, ka1 = Ms[r1] != 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c0 = nt * 8 + 2 * t, c1 = c0 + 1;
      const float Lc0 = Ls[c0], Lc1 = Ls[c1], Dc0 = Ds[c0], Dc1 = Ds[c1];
      const float p0 = ka0 ? exp2f(st[nt][0] * scale2 - Lc0) : 0.f, p1 = ka0 ? exp2f(st[nt][1] * scale2 - Lc1) : 0.f;
      const float p2 = ka1 ? exp2f(st[nt][2] * scale2 - Lc0) : 0.f, p3 = ka1 ? exp2f(st[nt][3] * scale2 - Lc1) : 0.f;
      st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
      dpt[nt][0] = p0 * (dpt[nt][0] - Dc0) * scale; dpt[nt][1] = p1 * (dpt[nt][1] - Dc1) * scale;
      dpt[nt][2] = p2 * (dpt[nt][2] - Dc0) * scale; dpt[nt][3] = p3 * (dpt[nt][3] - Dc1) * scale;
    }
    float dv[4][4], dk[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
      dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
    }
    mm_pv(dv, st, Gh, Gl, lane);            // dV = P^T dO
    mm_pv(dk, dpt, Qh, Ql, lane);           // dK = dS^T Q
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (r0 < S) {
        float* o = d_qkv + ((int64_t)b * S + r0) * ld + h * DK + nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(o + H) = make_float2(dk[nt][0], dk[nt][1]);
        *reinterpret_cast<float2*>(o + 2 * H) = make_float2(dv[nt][0], dv[nt][1]);
      }
      if (r1 < S) {
        float* o = d_qkv + ((int64_t)b * S + r1) * ld + h * DK + nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(o + H) = make_float2(dk[nt][2], dk[nt][3]);
        *reinterpret_cast<float2*>(o + 2 * H) = make_float2(dv[nt][2], dv[nt][3]);
      }
    }
  }
}

}  // namespace atc

// MVF_ATTN_TC: 0 = never, 1 (default) = when the caller allows reduced (2^-16) operand precision, 2 = always (tests)
static int attn_tc_mode() {
  const char* e = getenv("MVF_ATTN_TC");
  return e ? atoi(e) : 1;
}
bool attention_tc_ok(int dtype, int S, int dk, int H, const void* qkv, const void* other, bool allow_split) {
  const int mode = attn_tc_mode();
  if (mode == 0 || (mode == 1 && !allow_split)) return false;
  return dtype == MVF_F32 && S >= 1 && S <= atc::SP && dk == atc::DK && H % 4 == 0 && ((((uintptr_t)qkv) | ((uintptr_t)other)) & 15) == 0;
}
int attention_tc_fwd(int B, int S, int heads, const void* qkv, const float* keymask, void* ctx, float* lse, cudaStream_t st) {
  launch_k(atc::attn_tc_fwd_kernel, dim3(heads, B), 128, 0, st, S, heads * atc::DK, (const float*)qkv, keymask, (float*)ctx, lse);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
int attention_tc_bwd(int B, int S, int heads, const void* qkv, const float* keymask, const void* ctx, const float* lse,
                     const void* d_ctx, void* d_qkv, cudaStream_t st) {
  launch_k(atc::attn_tc_bwd_kernel, dim3(heads, B), 256, 0, st, S, heads * atc::DK, (const float*)qkv, keymask, (const float*)ctx, lse,
           (const float*)d_ctx, (float*)d_qkv);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace mvf

// PROTOTYPE (round 2 work item): the per-pair SCL kernel on the warp-level tensor cores.
// NOT on any default path: scl_fwd_bwd only takes it with MVF_SCL_TC=1, and it has not been run on a GPU yet -- it was
// written after this round's GPU budget was spent and is compile-checked only.  tests/test_gpu_scl.py carries a test
// that is skipped unless MVF_SCL_TC=1.
//
// Why: scl_pair_fused2_kernel (scl.cu) spends ~4 800 FMA warp-instructions per pair on the three T x T x D products
// and is instruction-bound at 19 % of the HBM copy bandwidth on a scaled batch.  The temporal-attention kernels
// (attention_tc.cu) showed that bf16 hi/lo operand splits on mma.sync.m16n8k16 (hi*hi + hi*lo + lo*hi, fp32
// accumulate) reproduce the fp64 formulation to ~2e-5; the same recipe here needs ~770 MMAs per pair.
//
// One CTA of four warps per video pair, T <= 32 frames per view (padded to 32), D % 32 == 0, D <= 256:
//   warps 0,1  direction 0: rows = view-0 frames i (16 each), columns = view-1 frames j;  S  = E0 E1^T
//   warps 2,3  direction 1: rows = view-1 frames j,           columns = view-0 frames i;  S' = E1 E0^T
// so every row statistic of scl.py:52-105 (partition sum with the masked-column extras `zext`, Gaussian label
// normaliser, KL terms, g = sum y r) is a reduction over the accumulator fragments of one warp (quad shuffles).  The
// per-direction gradient coefficients go through shared memory once, each warp adds the other direction's transposed
// tile to its own, and dE_rows = coef . E_cols / tau is the second MMA product with the coefficients re-used from the
// accumulators (C layout of two 8-column tiles = A layout of one 16-wide k-step).
#include <math.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace mvf {
namespace stc {

constexpr int TP = 32;   // padded frames per view
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
__device__ __forceinline__ float quad_add(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

static size_t smem_bytes(int D) {
  const size_t pitch = (size_t)D * 2 + 16;
  return 4 * TP * pitch                 // E0 hi, E0 lo, E1 hi, E1 lo
         + 2 * TP * (TP + 1) * 4        // per-direction gradient coefficients
         + 4 * TP * 4 + 64;             // steps, masks (two views), loss partials
}

__global__ void __launch_bounds__(128)
scl_pair_tc_kernel(const float* __restrict__ embs, const int64_t* __restrict__ seq_lens, const int64_t* __restrict__ steps,
                   const float* __restrict__ masks, int T, int D, float tau, float two_var, const float* __restrict__ Mptr,
                   const float* __restrict__ zext, float* __restrict__ c_out, float* __restrict__ loss_out,
                   float* __restrict__ d_embs) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t sm[];
  const int pitch = D * 2 + 16;
  const int MAT = TP * pitch;
  auto Eh = [&](int view) { return sm + (size_t)view * 2 * MAT; };          // view 0: hi | lo, view 1: hi | lo
  auto El = [&](int view) { return sm + (size_t)view * 2 * MAT + MAT; };
  float* coefx = reinterpret_cast<float*>(sm + 4 * MAT);     // [2][TP][TP + 1]
  float* st = coefx + 2 * TP * (TP + 1);                     // [2][TP] steps as float
  float* mk = st + 2 * TP;                                   // [2][TP] masks (0 beyond T)
  float* red = mk + 2 * TP;                                  // [4] loss partials
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int v = blockIdx.x;
  const int64_t row0g = (int64_t)v * 2 * T;                  // first global row of the pair (view 0), view 1 at + T

  // ---- stage the pair: fp32 -> bf16 hi / lo, rows >= T zero ----
  const int D4 = D >> 2;
  for (int idx = tid; idx < 2 * TP * D4; idx += blockDim.x) {
    const int view = idx / (TP * D4), rem = idx - view * TP * D4;
    const int row = rem / D4, c4 = rem - row * D4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < T) x = *reinterpret_cast<const float4*>(embs + (row0g + (int64_t)view * T + row) * D + 4 * c4);
    uint2 h, l;
    split2(x.x, x.y, h.x, l.x);
    split2(x.z, x.w, h.y, l.y);
    *reinterpret_cast<uint2*>(Eh(view) + row * pitch + c4 * 8) = h;
    *reinterpret_cast<uint2*>(El(view) + row * pitch + c4 * 8) = l;
  }
  for (int i = tid; i < 2 * TP; i += blockDim.x) {
    const int view = i / TP, a = i - view * TP;
    st[i] = a < T ? (float)steps[row0g + (int64_t)view * T + a] : 0.f;
    mk[i] = a < T ? masks[row0g + (int64_t)view * T + a] : 0.f;
  }
  __syncthreads();

  const int dir = warp >> 1, rb = (warp & 1) * 16;           // direction, first local row of this warp
  const int rv = dir, cv = 1 - dir;                          // row view, column view
  const float M = *Mptr, invM = 1.f / M;
  const float Lr = (float)seq_lens[v * 2 + rv], Lc = (float)seq_lens[v * 2 + cv];
  const float c_ex = LOG2E / tau, c_pw = -LOG2E / two_var;

  // ---- logits of this warp's 16 rows against the 32 columns ----
  float acc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  {
    const uint32_t a_off = (uint32_t)((rb + (lane & 15)) * pitch + (lane >> 4) * 16);
    const uint32_t b_off = (uint32_t)((lane & 7) * pitch + (lane >> 3) * 16);
    for (int kp = 0; kp < D / 32; ++kp) {                    // two 16-wide k-steps per iteration
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        ldsm_x4(smem_u32(Eh(rv)) + a_off + kp * 64 + ks * 32, ah[ks]);
        ldsm_x4(smem_u32(El(rv)) + a_off + kp * 64 + ks * 32, al[ks]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t bh[4], bl[4];                               // {b0, b1} of k-step 0, {b0, b1} of k-step 1
        ldsm_x4(smem_u32(Eh(cv)) + b_off + nt * 8 * pitch + kp * 64, bh);
        ldsm_x4(smem_u32(El(cv)) + b_off + nt * 8 * pitch + kp * 64, bl);
        mma3(acc[nt], ah[0], al[0], bh[0], bh[1], bl[0], bl[1]);
        mma3(acc[nt], ah[1], al[1], bh[2], bh[3], bl[2], bl[3]);
      }
    }
  }

  // ---- row statistics (rows a0 = rb + g and a1 = rb + g + 8 of view rv) ----
  const int a0 = rb + g, a1 = a0 + 8;
  const float* str = st + rv * TP;
  const float* stc = st + cv * TP;
  const float* mkr = mk + rv * TP;
  const float* mkc = mk + cv * TP;
  const bool rv0 = mkr[a0] != 0.f, rv1 = mkr[a1] != 0.f;     // masks are zero beyond T
  const float ar0 = __fdiv_rn(str[a0], Lr), ar1 = __fdiv_rn(str[a1], Lr);   // scl.py:62, torch float32 op order
  float pw[4][4];                                            // log2 of the label weights (-inf: masked pair)
  float den0 = 0.f, den1 = 0.f, zp0 = 0.f, zp1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int b = nt * 8 + 2 * t + e;
      const bool cvd = mkc[b] != 0.f;
      const float sc = stc[b];
      const float d0 = fabsf(__fsub_rn(__fmul_rn(ar0, Lc), sc)), d1 = fabsf(__fsub_rn(__fmul_rn(ar1, Lc), sc));
      pw[nt][e] = (rv0 && cvd) ? d0 * d0 * c_pw : -INFINITY;
      pw[nt][2 + e] = (rv1 && cvd) ? d1 * d1 * c_pw : -INFINITY;
      acc[nt][e] = exp2f(acc[nt][e] * c_ex);                 // e^{l}
      acc[nt][2 + e] = exp2f(acc[nt][2 + e] * c_ex);
      den0 += exp2f(pw[nt][e]);
      den1 += exp2f(pw[nt][2 + e]);
      if (rv0 && cvd) zp0 += acc[nt][e];
      if (rv1 && cvd) zp1 += acc[nt][2 + e];
    }
  }
  den0 = quad_add(den0); den1 = quad_add(den1);
  zp0 = quad_add(zp0); zp1 = quad_add(zp1);
  const int64_t gr0 = row0g + (int64_t)rv * T + a0, gr1 = gr0 + 8;
  const float Z0 = zp0 + (a0 < T ? zext[gr0] : 0.f), Z1 = zp1 + (a1 < T ? zext[gr1] : 0.f);
  const bool live0 = rv0 && Z0 > 0.f, live1 = rv1 && Z1 > 0.f;
  const float iZ0 = live0 ? 1.f / Z0 : 0.f, iZ1 = live1 ? 1.f / Z1 : 0.f;
  const float ld0 = den0 > 0.f ? log2f(den0) : INFINITY, ld1 = den1 > 0.f ? log2f(den1) : INFINITY;
  float g0 = 0.f, g1 = 0.f, loss = 0.f;
  float yr[4][4];                                            // y r per element (needed again for the coefficients)
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool lv = e < 2 ? live0 : live1;
      const float y = lv ? exp2f(pw[nt][e] - (e < 2 ? ld0 : ld1)) : 0.f;
      const float p = acc[nt][e] * (e < 2 ? iZ0 : iZ1);
      float r = 0.f;
      if (y > 1e-30f) {                                      // as in scl_pair_fused2_kernel: terms below 1e-30 are dropped
        const float q = p + 1e-6f;
        loss += y * (__logf(y) - __logf(q));
        r = y * __fdividef(p, q);
        if (e < 2) g0 += r; else g1 += r;
      }
      yr[nt][e] = r;
      acc[nt][e] = p;                                        // keep p
    }
  }
  g0 = quad_add(g0);
  g1 = quad_add(g1);
  if (t == 0) {
    if (a0 < T) c_out[gr0] = live0 ? g0 * iZ0 * invM : 0.f;
    if (a1 < T) c_out[gr1] = live1 ? g1 * iZ1 * invM : 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  if (lane == 0) red[warp] = loss;

  // ---- gradient coefficients of this direction -> shared memory ----
  float* mine = coefx + dir * TP * (TP + 1);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int b = nt * 8 + 2 * t + (e & 1);
      const bool cvd = mkc[b] != 0.f;
      const bool lv = e < 2 ? live0 : live1;
      const float cf = (lv && cvd) ? (acc[nt][e] * (e < 2 ? g0 : g1) - yr[nt][e]) * invM : 0.f;
      acc[nt][e] = cf;
      mine[(e < 2 ? a0 : a1) * (TP + 1) + b] = cf;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const float tsum = (red[0] + red[1]) + (red[2] + red[3]);
    if (tsum != 0.f) atomicAdd(loss_out, tsum * invM);
  }
  if (d_embs == nullptr) return;

  // ---- total coefficient in this warp's orientation: own + the other direction's transposed tile ----
  const float* other = coefx + (1 - dir) * TP * (TP + 1);
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int b = nt * 8 + 2 * t + (e & 1);
      acc[nt][e] += other[b * (TP + 1) + (e < 2 ? a0 : a1)];
    }
  }

  // ---- dE_rows = coef . E_cols / tau, 16 channels (two 8-wide tiles) at a time ----
  uint32_t ch[2][4], cl[2][4];                               // A fragments of the coefficient tile, two k-steps of 16 columns
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    split2(acc[2 * kk][0], acc[2 * kk][1], ch[kk][0], cl[kk][0]);
    split2(acc[2 * kk][2], acc[2 * kk][3], ch[kk][1], cl[kk][1]);
    split2(acc[2 * kk + 1][0], acc[2 * kk + 1][1], ch[kk][2], cl[kk][2]);
    split2(acc[2 * kk + 1][2], acc[2 * kk + 1][3], ch[kk][3], cl[kk][3]);
  }
  const float inv_tau = 1.f / tau;
  const uint32_t v_off = (uint32_t)((lane & 15) * pitch + (lane >> 4) * 16);
  float* out0 = d_embs + gr0 * D;
  float* out1 = d_embs + gr1 * D;
  for (int cp = 0; cp < D / 16; ++cp) {
    float o[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) o[u][0] = o[u][1] = o[u][2] = o[u][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t bh[4], bl[4];                                 // {b0, b1} of channel tile 2cp, {b0, b1} of tile 2cp + 1
      ldsm_x4_trans(smem_u32(Eh(cv)) + v_off + kk * 16 * pitch + cp * 32, bh);
      ldsm_x4_trans(smem_u32(El(cv)) + v_off + kk * 16 * pitch + cp * 32, bl);
      mma3(o[0], ch[kk], cl[kk], bh[0], bh[1], bl[0], bl[1]);
      mma3(o[1], ch[kk], cl[kk], bh[2], bh[3], bl[2], bl[3]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int c = cp * 16 + u * 8 + 2 * t;
      if (a0 < T) *reinterpret_cast<float2*>(out0 + c) = make_float2(o[u][0] * inv_tau, o[u][1] * inv_tau);
      if (a1 < T) *reinterpret_cast<float2*>(out1 + c) = make_float2(o[u][2] * inv_tau, o[u][3] * inv_tau);
    }
  }
}

}  // namespace stc

bool scl_pair_tc_enabled(int T, int D) {
  const char* e = getenv("MVF_SCL_TC");
  if (!e || atoi(e) == 0) return false;
  return T >= 1 && T <= stc::TP && D % 32 == 0 && D <= 256 && stc::smem_bytes(D) <= 200 * 1024;
}

int scl_pair_tc(const float* embs, const int64_t* seq_lens, const int64_t* steps, const float* masks, int Bv, int T, int D,
                float tau, float two_var, const float* Mptr, const float* zext, float* c_out, float* loss_out, float* d_embs,
                cudaStream_t st) {
  const size_t smem = stc::smem_bytes(D);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    MVF_CHECK_CUDA(cudaFuncSetAttribute(stc::scl_pair_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch_k(stc::scl_pair_tc_kernel, Bv, 128, smem, st, embs, seq_lens, steps, masks, T, D, tau, two_var, Mptr, zext, c_out, loss_out,
           d_embs);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace mvf

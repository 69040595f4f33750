// Sequence Contrastive Loss (algos/scl.py:52-105), NEGATIVE_TYPE single_noself / batch_noself: forward + gradient on the
// warp-level tensor cores, T <= 256 frames per view, D <= 256 channels.
//
// Every product of the loss is a small matrix product between the two views of ONE video pair (logits S = E0 E1^T, gradients
// dE0 = C E1, dE1 = C^T E0 with C = dloss/dS) or between a row block and the batch's masked frames (the scl.py:80 quirk: every
// masked frame enters every valid row's partition sum with weight 1e-6).  They run as mma.sync.m16n8k16 on bf16 hi / lo operand
// splits (hi*hi + hi*lo + lo*hi, fp32 accumulate: ~2^-17 per product, measured ~4e-6 on the gradient), with the softmax /
// Gaussian-label / KL arithmetic done on the accumulator fragments: nothing of size T x T or N x N reaches shared or global
// memory, and a pair's embeddings are read from HBM once and its gradient written once.
//
// Work unit = a warp with 16 rows (frames) of one view of one pair; its columns are the partner view's frames, 32 per tile:
//   pass A   S tile -> e^{l}; partition sum Z over the valid partner columns (+ the extras of the cross pass), label normaliser
//   pass B   S tile -> y, p, r: loss, g = sum y r, c = g / (Z M)                     (row statistics -> partner rows' warps)
//   phase 2  S tile -> coefficient of BOTH view directions at (row, column) from the row's and the column's statistics;
//            the accumulator fragments of two 8-column tiles are the A fragments of one 16-wide k-step of dE += C . E_cols
// scl_pair_mma_kernel<.., BOTH, KEEP>
//   BOTH + KEEP  T <= 32: one CTA (4 warps) holds both views; e^{l} stays in registers through all passes, the partner operand
//                is the other view's panel (no second staging), row statistics cross through shared memory, every 16-channel
//                slice of the gradient leaves as soon as its two k-steps are done (no D-wide accumulator);
//   BOTH         T <= 96 (D <= 128) / T <= 64 (D <= 256), batch >= the SM count: the same with S recomputed per pass;
//   cluster      the row blocks of a pair are spread over a thread-block cluster (2 views x 1/2/4 chunks, more chunks when the
//                batch alone cannot fill the machine); each CTA stages the partner view once when it fits (else per pass in
//                multiples of 32 columns), and the row statistics cross through global memory + one cluster barrier.
// scl_cross_mma_kernel<.., GRAD>: the terms that couple a row to frames outside its pair, rows and columns taken from the
//   compacted valid / masked lists of scl_prep (every job of a launch is one (row list, column list, weights) triple):
//     quirk  (scl.py:80)    rows valid, cols masked, weight 1e-6:  Z extras; dE_r += c_r sum_k x_rk e_k / tau
//                           rows masked, cols valid:               dE_k += sum_r c_r x_rk e_r / tau
//     batch_noself (74-79)  rows valid, cols valid of OTHER videos, weight 1: the same three terms
//   64 rows per CTA (a warp per 16), the column tiles strided over blockIdx.y, sums with atomicAdd / red.global.add.v2.f32.
#include <math.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "scl_ws.cuh"

namespace mvf {
namespace smma {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int CT = 32;   // columns per tile

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
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpa(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void red_add2(float* p, float x, float y) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}

// Cooperative staging of `nrows` embedding rows (rowof(r) = global row, < 0: zeros) as bf16 hi / lo panels of `pitch` bytes.
// A thread keeps its 16-byte column and walks rows (the usual case: the block size is a multiple of the row's float4 count);
// eight loads per thread are in flight before the first conversion.
template <typename F>
__device__ __forceinline__ void stage_rows(uint8_t* hi, uint8_t* lo, int pitch, int nrows, int D, int Dp,
                                           const float* __restrict__ embs, F rowof) {
  const int D4p = Dp >> 2, nthr = blockDim.x;
  const int rstep = nthr / D4p;
  if (rstep * D4p == nthr) {
    const int r0 = threadIdx.x / D4p, c4 = threadIdx.x - r0 * D4p;
    const bool cin = 4 * c4 < D;
    const float* src = embs + 4 * c4;
    uint8_t* dh = hi + c4 * 8;
    uint8_t* dl = lo + c4 * 8;
    for (int rbase = r0; rbase < nrows; rbase += 8 * rstep) {
      float4 x[8];
      unsigned valid = 0;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = rbase + u * rstep;
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < nrows) {
          const int64_t gr = rowof(row);
          if (gr >= 0 && cin) {
            x[u] = __ldg(reinterpret_cast<const float4*>(src + gr * D));
            valid |= 1u << u;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = rbase + u * rstep;
        if (row < nrows) {
          uint2 h = make_uint2(0u, 0u), l = make_uint2(0u, 0u);
          if (valid & (1u << u)) {                            // padded rows are stored as zeros without the conversion
            split2(x[u].x, x[u].y, h.x, l.x);
            split2(x[u].z, x[u].w, h.y, l.y);
          }
          *reinterpret_cast<uint2*>(dh + row * pitch) = h;
          *reinterpret_cast<uint2*>(dl + row * pitch) = l;
        }
      }
    }
    return;
  }
  const int total = nrows * D4p;
  for (int base = threadIdx.x; base < total; base += 8 * nthr) {
    float4 x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * nthr;
      x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < total) {
        const int row = idx / D4p, c4 = idx - row * D4p;
        const int64_t gr = rowof(row);
        if (gr >= 0 && 4 * c4 < D) x[u] = __ldg(reinterpret_cast<const float4*>(embs + gr * D) + c4);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * nthr;
      if (idx < total) {
        const int row = idx / D4p, c4 = idx - row * D4p;
        uint2 h, l;
        split2(x[u].x, x[u].y, h.x, l.x);
        split2(x[u].z, x[u].w, h.y, l.y);
        *reinterpret_cast<uint2*>(hi + row * pitch + c4 * 8) = h;
        *reinterpret_cast<uint2*>(lo + row * pitch + c4 * 8) = l;
      }
    }
  }
}

// 16 x 32 tile of E_rows E_cols^T (the first ntv 8-column tiles; the others stay zero).  pa_*: shared address of the warp's
// row block + its ldmatrix lane offset ((lane & 15) * pitch + (lane >> 4) * 16); cb_*: tile base + (lane & 7) * pitch +
// (lane >> 3) * 16.
__device__ __forceinline__ void s_tile(float (&acc)[4][4], uint32_t pa_hi, uint32_t pa_lo, uint32_t cb_hi, uint32_t cb_lo,
                                       int pitch, int Dp, int ntv) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  for (int kp = 0; kp < (Dp >> 5); ++kp) {                   // two 16-wide k-steps per iteration
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      ldsm_x4(pa_hi + kp * 64 + ks * 32, ah[ks]);
      ldsm_x4(pa_lo + kp * 64 + ks * 32, al[ks]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (nt < ntv) {
        uint32_t bh[4], bl[4];                               // {b0, b1} of k-step 0, {b0, b1} of k-step 1
        ldsm_x4(cb_hi + nt * 8 * pitch + kp * 64, bh);
        ldsm_x4(cb_lo + nt * 8 * pitch + kp * 64, bl);
        mma3(acc[nt], ah[0], al[0], bh[0], bh[1], bl[0], bl[1]);
        mma3(acc[nt], ah[1], al[1], bh[2], bh[3], bl[2], bl[3]);
      }
    }
  }
}

// The C fragments of two 8-column tiles are the A fragments of one 16-wide k-step: coefficient tile -> bf16 hi / lo A operands.
__device__ __forceinline__ void coef_frags(const float (&cf)[4][4], uint32_t (&ch)[2][4], uint32_t (&cl)[2][4]) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    split2(cf[2 * kk][0], cf[2 * kk][1], ch[kk][0], cl[kk][0]);
    split2(cf[2 * kk][2], cf[2 * kk][3], ch[kk][1], cl[kk][1]);
    split2(cf[2 * kk + 1][0], cf[2 * kk + 1][1], ch[kk][2], cl[kk][2]);
    split2(cf[2 * kk + 1][2], cf[2 * kk + 1][3], ch[kk][3], cl[kk][3]);
  }
}
// channels [16 cp, 16 cp + 16) of o[16 x Dp] += cf[16 x 32] . E_cols[32 x Dp].  tv_*: tile base + (lane & 15) * pitch +
// (lane >> 4) * 16; ksv: k-steps (16 columns each) that hold real columns.
__device__ __forceinline__ void de_chunk(float (&o0)[4], float (&o1)[4], const uint32_t (&ch)[2][4], const uint32_t (&cl)[2][4],
                                         uint32_t tv_hi, uint32_t tv_lo, int pitch, int cp, int ksv) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    if (kk < ksv) {
      uint32_t bh[4], bl[4];                                 // {b0, b1} of channel tile 2cp, {b0, b1} of tile 2cp + 1
      ldsm_x4_trans(tv_hi + kk * 16 * pitch + cp * 32, bh);
      ldsm_x4_trans(tv_lo + kk * 16 * pitch + cp * 32, bl);
      mma3(o0, ch[kk], cl[kk], bh[0], bh[1], bl[0], bl[1]);
      mma3(o1, ch[kk], cl[kk], bh[2], bh[3], bl[2], bl[3]);
    }
  }
}
template <int NTD>
__device__ __forceinline__ void de_tile(float (&o)[NTD][4], const float (&cf)[4][4], uint32_t tv_hi, uint32_t tv_lo, int pitch,
                                        int Dp, int ksv) {
  uint32_t ch[2][4], cl[2][4];
  coef_frags(cf, ch, cl);
#pragma unroll
  for (int cp = 0; cp < NTD / 2; ++cp)
    if (cp * 16 < Dp) de_chunk(o[2 * cp], o[2 * cp + 1], ch, cl, tv_hi, tv_lo, pitch, cp, ksv);
}

struct PairArgs {
  const float* embs;
  const int64_t* seq_lens;
  const int64_t* steps;
  const float* masks;
  int T, D;
  float tau, two_var;
  SclWs w;
  int use_zext;       // 1: add w.zext (extras of the cross pass: masked frames / batch negatives) to Z
  float* loss_out;
  float* d_embs;      // null: loss only
  int chunks;         // cluster mode: CTAs per view
  int Wc;             // cluster mode: row blocks (compute warps) per CTA; further warps only help staging
  int CS;             // cluster mode: partner columns resident per staging (multiple of 32)
};

static __host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// BOTH: one CTA holds both views of the pair (warp = view * nbv + row block); else a cluster of 2 * chunks CTAs.
// KEEP (BOTH, T <= 32): e^{l} of the single partner tile stays in registers through every pass.
// NTVC (KEEP only): ceil(T / 8) as a compile-time constant, so that the guards of the padded 8-column tiles fold away.
template <int NTD, bool BOTH, bool KEEP, int NTHR, int MINB, int NTVC = 0>
__global__ void __launch_bounds__(NTHR, MINB) scl_pair_mma_kernel(const PairArgs A) {
  pdl_entry();
  extern __shared__ __align__(128) uint8_t sm[];
  const int T = A.T, D = A.D;
  const int Tp = round_up(T, CT), Dp = round_up(D, 32), pitch = Dp * 2 + 16, nbv = (T + 15) >> 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int NW = blockDim.x >> 5;
  int v, view, rbk, rb0;
  bool active;
  if (BOTH) {
    v = blockIdx.x; view = warp / nbv; rbk = warp - view * nbv; rb0 = 0;
    active = warp < 2 * nbv;
    if (!active) { view = 0; rbk = 0; }
  } else {
    const int CL = 2 * A.chunks, rank = blockIdx.x % CL;
    v = blockIdx.x / CL; view = rank / A.chunks; rb0 = (rank % A.chunks) * A.Wc; rbk = rb0 + warp;
    active = warp < A.Wc && rbk < nbv;
  }
  const int cv = 1 - view;
  const int64_t row0g = (int64_t)v * 2 * T;

  // ---- shared memory ----
  const int PR = BOTH ? 2 * Tp : A.Wc * 16;                  // panel rows
  const int CS = BOTH ? Tp : A.CS;
  uint8_t* pan_hi = sm;
  uint8_t* pan_lo = pan_hi + (size_t)PR * pitch;
  uint8_t* col_hi = BOTH ? pan_hi : pan_lo + (size_t)PR * pitch;     // BOTH: the partner operand is the other view's panel
  uint8_t* col_lo = BOTH ? pan_lo : col_hi + (size_t)CS * pitch;
  float* fm = reinterpret_cast<float*>(BOTH ? pan_lo + (size_t)PR * pitch : col_lo + (size_t)CS * pitch);
  const int NC = BOTH ? 2 * Tp : Tp;                         // column metadata entries
  float* cst = fm;            // steps; 1e30 for a masked / padded frame (its label weight becomes exactly 0)
  float* cmk = cst + NC;      // 1 valid, 0 masked / padded
  float* car = cmk + NC;      // fl(s / L) of the column's own view
  float* cZ = car + NC;       // 1 / Z (0: row not live)
  float* cg = cZ + NC;        // g
  float* cl2 = cg + NC;       // log2 den (+inf: row not live)
  float* red = cl2 + NC;      // [32] loss partials
  const int cmb = BOTH ? cv * Tp : 0;                        // first metadata entry of this warp's partner view

  // ---- stage: own rows (BOTH: both views), partner metadata ----
  if (BOTH) {
    stage_rows(pan_hi, pan_lo, pitch, 2 * Tp, D, Dp, A.embs, [&](int r) -> int64_t {
      const int vw = r >= Tp, fr = r - vw * Tp;
      return fr < T ? row0g + (int64_t)vw * T + fr : -1;
    });
    for (int i = tid; i < 2 * Tp; i += blockDim.x) {
      const int vw = i >= Tp, fr = i - vw * Tp;
      float s = 0.f, m = 0.f;
      if (fr < T) { s = (float)A.steps[row0g + (int64_t)vw * T + fr]; m = A.masks[row0g + (int64_t)vw * T + fr]; }
      cst[i] = m != 0.f ? s : 1e30f; cmk[i] = m != 0.f ? 1.f : 0.f;
      car[i] = __fdiv_rn(s, (float)A.seq_lens[v * 2 + vw]);
    }
  } else {
    stage_rows(pan_hi, pan_lo, pitch, PR, D, Dp, A.embs, [&](int r) -> int64_t {
      return rb0 * 16 + r < T ? row0g + (int64_t)view * T + rb0 * 16 + r : -1;
    });
    const float Lcv = (float)A.seq_lens[v * 2 + cv];
    for (int i = tid; i < Tp; i += blockDim.x) {
      float s = 0.f, m = 0.f;
      if (i < T) { s = (float)A.steps[row0g + (int64_t)cv * T + i]; m = A.masks[row0g + (int64_t)cv * T + i]; }
      cst[i] = m != 0.f ? s : 1e30f; cmk[i] = m != 0.f ? 1.f : 0.f;
      car[i] = __fdiv_rn(s, Lcv);
    }
  }
  const bool resident = BOTH || CS >= Tp;
  auto stage_cols = [&](int s0) {
    stage_rows(col_hi, col_lo, pitch, min(CS, Tp - s0), D, Dp, A.embs, [&](int r) -> int64_t {
      return s0 + r < T ? row0g + (int64_t)cv * T + s0 + r : -1;
    });
  };
  if (!BOTH && resident) stage_cols(0);
  __syncthreads();

  // ---- this thread's two rows: a[0] = 16 rbk + g, a[1] = a[0] + 8 ----
  const float M = *A.w.M, invM = 1.f / M;
  const float Lr = (float)A.seq_lens[v * 2 + view], Lc = (float)A.seq_lens[v * 2 + cv];
  const float c_ex = LOG2E / A.tau, c_pw = -LOG2E / A.two_var, inv_tau = 1.f / A.tau;
  int a[2];
  int64_t grow[2];
  bool rv[2];
  float str[2], arr[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    a[h] = rbk * 16 + g + 8 * h;
    grow[h] = row0g + (int64_t)view * T + a[h];
    const bool in = active && a[h] < T;
    rv[h] = in && A.masks[grow[h]] != 0.f;
    str[h] = in ? (float)A.steps[grow[h]] : 0.f;
    arr[h] = __fdiv_rn(str[h], Lr);                          // scl.py:62, torch float32 op order
  }
  const int prow = BOTH ? view * Tp + rbk * 16 : warp * 16;  // this warp's first panel row
  const uint32_t a_off = (uint32_t)((prow + (lane & 15)) * pitch + (lane >> 4) * 16);
  const uint32_t pa_hi = smem_u32(pan_hi) + a_off, pa_lo = smem_u32(pan_lo) + a_off;
  const uint32_t b_off = (uint32_t)((lane & 7) * pitch + (lane >> 3) * 16);
  const uint32_t v_off = (uint32_t)((lane & 15) * pitch + (lane >> 4) * 16);
  const uint32_t colbase_hi = smem_u32(col_hi) + (BOTH ? (uint32_t)(cv * Tp * pitch) : 0u);
  const uint32_t colbase_lo = smem_u32(col_lo) + (BOTH ? (uint32_t)(cv * Tp * pitch) : 0u);
  const float* mcst = cst + cmb + 2 * t;                     // this thread's columns: mcst[c0 + 8 nt + (e & 1)]
  const float* mcmk = cmk + cmb + 2 * t;
  const float* mcar = car + cmb + 2 * t;
  const float* mcZ = cZ + cmb + 2 * t;
  const float* mcg = cg + cmb + 2 * t;
  const float* mcl2 = cl2 + cmb + 2 * t;

  // every partner tile of one pass: body(c0, tile_hi, tile_lo, ntv); all threads of the CTA walk the staging barriers
  auto partner_pass = [&](auto&& body) {
    if constexpr (KEEP) {
      if (active) body(0, colbase_hi, colbase_lo, NTVC ? NTVC : min(4, (T + 7) >> 3));
    } else {
      for (int s0 = 0; s0 < Tp; s0 += CS) {
        if (!resident) {
          __syncthreads();
          stage_cols(s0);
          __syncthreads();
        }
        if (active)
          for (int c0 = s0; c0 < min(s0 + CS, Tp); c0 += CT)
            body(c0, colbase_hi + (uint32_t)((c0 - s0) * pitch), colbase_lo + (uint32_t)((c0 - s0) * pitch), min(4, (T - c0 + 7) >> 3));
      }
    }
  };
  // log2 of the Gaussian label weight of (row h, column index j of the thread); -inf when the column is masked
  auto label_pw = [&](int h, int j) -> float {
    const float d = __fsub_rn(__fmul_rn(arr[h], Lc), mcst[j]);
    return d * d * c_pw;
  };

  float ex[4][4];                                            // KEEP: e^{l} of the pair's tile, kept through every pass
  float den[2] = {0.f, 0.f}, zp[2] = {0.f, 0.f};

  // ---- pass A: partition sums and label normalisers ----
  partner_pass([&](int c0, uint32_t th, uint32_t tl, int ntv) {
    float acc[4][4];
    s_tile(acc, pa_hi, pa_lo, th + b_off, tl + b_off, pitch, Dp, ntv);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (nt < ntv) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h = e >> 1, j = c0 + nt * 8 + (e & 1);
          const float x = ex2(acc[nt][e] * c_ex);
          den[h] += ex2(label_pw(h, j));
          zp[h] = fmaf(x, mcmk[j], zp[h]);
          if (KEEP) ex[nt][e] = x;
        }
      } else if (KEEP) {
        ex[nt][0] = ex[nt][1] = ex[nt][2] = ex[nt][3] = 0.f;
      }
    }
  });
  float iZ[2], l2d[2], gs[2] = {0.f, 0.f}, cr[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    den[h] = quad_add(den[h]);
    float Z = quad_add(zp[h]);
    if (A.use_zext && active && a[h] < T) Z += A.w.zext[grow[h]];
    const bool live = rv[h] && Z > 0.f;
    iZ[h] = live ? 1.f / Z : 0.f;
    l2d[h] = (live && den[h] > 0.f) ? log2f(den[h]) : INFINITY;      // a row that is not live gets y = 0 everywhere
  }

  // ---- pass B: loss (in log2 units) and g ----
  float loss2 = 0.f;
  auto pass_b = [&](int c0, const float (&xs)[4][4], int ntv) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (nt < ntv) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h = e >> 1, j = c0 + nt * 8 + (e & 1);
          const float y = ex2(label_pw(h, j) - l2d[h]);
          const float p = xs[nt][e] * iZ[h], q = p + 1e-6f;
          loss2 = fmaf(y, lg2(fmaxf(y, 1e-37f)) - lg2(q), loss2);    // y log y -> 0 for y -> 0 (no 0 * inf)
          gs[h] = fmaf(y * p, rcpa(q), gs[h]);
        }
      }
    }
  };
  if (KEEP) {
    if (active) pass_b(0, ex, NTVC ? NTVC : min(4, (T + 7) >> 3));
  } else {
    partner_pass([&](int c0, uint32_t th, uint32_t tl, int ntv) {
      float acc[4][4];
      s_tile(acc, pa_hi, pa_lo, th + b_off, tl + b_off, pitch, Dp, ntv);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = ex2(acc[nt][e] * c_ex);
      pass_b(c0, acc, ntv);
    });
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    gs[h] = quad_add(gs[h]);
    cr[h] = gs[h] * iZ[h] * invM;
  }
  const float loss = warp_sum(loss2) * 0.6931471805599453f;
  if (lane == 0) red[warp] = active ? loss : 0.f;

  // ---- row statistics to the partner rows' warps ----
  if (active && t == 0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (a[h] < T) A.w.c[grow[h]] = cr[h];
      if (BOTH) {
        const int i = view * Tp + a[h];
        cZ[i] = iZ[h]; cg[i] = gs[h]; cl2[i] = l2d[h];
      } else if (a[h] < T) {
        A.w.Z[grow[h]] = iZ[h]; A.w.g[grow[h]] = gs[h]; A.w.den[grow[h]] = l2d[h];
      }
    }
  }
  if (BOTH) {
    __syncthreads();
  } else {
    __threadfence();
    cluster_sync_all();
    for (int i = tid; i < Tp; i += blockDim.x) {
      const bool in = i < T;
      const int64_t r = row0g + (int64_t)cv * T + i;
      cZ[i] = in ? __ldcg(A.w.Z + r) : 0.f;
      cg[i] = in ? __ldcg(A.w.g + r) : 0.f;
      cl2[i] = in ? __ldcg(A.w.den + r) : INFINITY;
    }
    __syncthreads();
  }
  if (tid == 0) {
    float tsum = 0.f;
    for (int k = 0; k < NW; ++k) tsum += red[k];
    if (tsum != 0.f) atomicAdd(A.loss_out, tsum * invM);
  }
  if (A.d_embs == nullptr) return;

  // ---- phase 2: dE_rows = sum_cols coef . e_col / tau ----
  // e^{l} -> dloss/dl of both view directions, in place.  A row / column that is not live has 1/Z = 0 and log2 den = +inf,
  // so its direction contributes p g - y r = 0 by itself; a masked column of a live row is cut by the final select.
  auto coef_tile = [&](int c0, float (&xs)[4][4], int ntv) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int h = e >> 1, j = c0 + nt * 8 + (e & 1);
        float cf = 0.f;
        if (nt < ntv) {
          const float x = xs[nt][e];
          const float p0 = x * iZ[h];
          const float y0 = ex2(label_pw(h, j) - l2d[h]);
          const float t0 = fmaf(p0, gs[h], -y0 * p0 * rcpa(p0 + 1e-6f));
          const float p1 = x * mcZ[j];
          const float d1 = __fsub_rn(__fmul_rn(mcar[j], Lr), str[h]);
          const float y1 = ex2(d1 * d1 * c_pw - mcl2[j]);
          const float t1 = fmaf(p1, mcg[j], -y1 * p1 * rcpa(p1 + 1e-6f));
          cf = (rv[h] && mcmk[j] != 0.f) ? (t0 + t1) * invM : 0.f;
        }
        xs[nt][e] = cf;
      }
    }
  };
  if constexpr (KEEP) {
    // one partner tile: every 16-channel slice of the gradient is complete after two k-steps and leaves at once
    if (active) {
      const int ntv = NTVC ? NTVC : min(4, (T + 7) >> 3);
      coef_tile(0, ex, ntv);
      uint32_t ch[2][4], cl[2][4];
      coef_frags(ex, ch, cl);
      float* out0 = A.d_embs + grow[0] * D + 2 * t;
      float* out1 = A.d_embs + grow[1] * D + 2 * t;
      for (int cp = 0; cp < (Dp >> 4); ++cp) {
        float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
        de_chunk(o0, o1, ch, cl, colbase_hi + v_off, colbase_lo + v_off, pitch, cp, (ntv + 1) >> 1);
        const int c = cp * 16 + 2 * t;
        if (a[0] < T) {
          if (c < D) *reinterpret_cast<float2*>(out0 + cp * 16) = make_float2(o0[0] * inv_tau, o0[1] * inv_tau);
          if (c + 8 < D) *reinterpret_cast<float2*>(out0 + cp * 16 + 8) = make_float2(o1[0] * inv_tau, o1[1] * inv_tau);
        }
        if (a[1] < T) {
          if (c < D) *reinterpret_cast<float2*>(out1 + cp * 16) = make_float2(o0[2] * inv_tau, o0[3] * inv_tau);
          if (c + 8 < D) *reinterpret_cast<float2*>(out1 + cp * 16 + 8) = make_float2(o1[2] * inv_tau, o1[3] * inv_tau);
        }
      }
    }
  } else {
    float o[NTD][4];
#pragma unroll
    for (int n = 0; n < NTD; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    partner_pass([&](int c0, uint32_t th, uint32_t tl, int ntv) {
      float acc[4][4];
      s_tile(acc, pa_hi, pa_lo, th + b_off, tl + b_off, pitch, Dp, ntv);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = ex2(acc[nt][e] * c_ex);
      coef_tile(c0, acc, ntv);
      de_tile<NTD>(o, acc, th + v_off, tl + v_off, pitch, Dp, (ntv + 1) >> 1);
    });
    if (active) {
#pragma unroll
      for (int n = 0; n < NTD; ++n) {
        const int c = n * 8 + 2 * t;
        if (c < D) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (a[h] < T)
              *reinterpret_cast<float2*>(A.d_embs + grow[h] * D + c) = make_float2(o[n][2 * h] * inv_tau, o[n][2 * h + 1] * inv_tau);
        }
      }
    }
  }
}

// Cross terms: rows r of a row list against columns k of a column list,
//   x_rk = w0 [vid(r) != vid(k) or no exclusion] exp(<e_r, e_k> / tau)
//   GRAD = false:  sum_out[r] += sum_k x_rk                                   (partition-sum extras)
//   GRAD = true:   vec_out[r, :] += rc_r sum_k cc_k x_rk e_k / tau            (rc / cc null: 1)
// blockIdx.z = job, blockIdx.x = 64-row blocks (grid-stride), blockIdx.y = column tiles (strided).
template <int NTD, bool GRAD>
__global__ void __launch_bounds__(128) scl_cross_mma_kernel(const float* __restrict__ embs, int D, int T2, float tau,
                                                            const SclCrossJobs J, float* __restrict__ sum_out,
                                                            float* __restrict__ vec_out) {
  pdl_entry();
  extern __shared__ __align__(128) uint8_t sm[];
  const SclCrossJob& jb = J.job[blockIdx.z];
  const int nr = *jb.row_cnt, nc = *jb.col_cnt;
  if (nr == 0 || nc == 0 || (int)blockIdx.y * CT >= nc) return;
  const int Dp = round_up(D, 32), pitch = Dp * 2 + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  uint8_t* pan_hi = sm;
  uint8_t* pan_lo = pan_hi + (size_t)64 * pitch;
  uint8_t* tb_hi = pan_lo + (size_t)64 * pitch;
  uint8_t* tb_lo = tb_hi + (size_t)CT * pitch;
  float* cc = reinterpret_cast<float*>(tb_lo + (size_t)CT * pitch);   // [32] column weight (0 beyond the list)
  int* cvid = reinterpret_cast<int*>(cc + CT);                        // [32] column video
  const float c_ex = LOG2E / tau, inv_tau = 1.f / tau;
  const uint32_t a_off = (uint32_t)((warp * 16 + (lane & 15)) * pitch + (lane >> 4) * 16);
  const uint32_t pa_hi = smem_u32(pan_hi) + a_off, pa_lo = smem_u32(pan_lo) + a_off;
  const uint32_t b_off = (uint32_t)((lane & 7) * pitch + (lane >> 3) * 16);
  const uint32_t v_off = (uint32_t)((lane & 15) * pitch + (lane >> 4) * 16);
  const uint32_t tbase_hi = smem_u32(tb_hi), tbase_lo = smem_u32(tb_lo);
  for (int rb = blockIdx.x; rb * 64 < nr; rb += gridDim.x) {
    __syncthreads();
    stage_rows(pan_hi, pan_lo, pitch, 64, D, Dp, embs, [&](int r) -> int64_t { return rb * 64 + r < nr ? (int64_t)jb.row_idx[rb * 64 + r] : -1; });
    const bool wact = rb * 64 + warp * 16 < nr;
    int row[2], rvid[2];
    float rw[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int slot = rb * 64 + warp * 16 + g + 8 * h;
      row[h] = slot < nr ? jb.row_idx[slot] : -1;
      rvid[h] = row[h] >= 0 ? row[h] / T2 : -1;
      rw[h] = row[h] < 0 ? 0.f : (GRAD && jb.rc) ? jb.w0 * jb.rc[row[h]] : jb.w0;
    }
    float o[GRAD ? NTD : 1][4];
#pragma unroll
    for (int n = 0; n < (GRAD ? NTD : 1); ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float rsum[2] = {0.f, 0.f};
    for (int c0 = blockIdx.y * CT; c0 < nc; c0 += CT * gridDim.y) {
      __syncthreads();
      stage_rows(tb_hi, tb_lo, pitch, CT, D, Dp, embs, [&](int r) -> int64_t { return c0 + r < nc ? (int64_t)jb.col_idx[c0 + r] : -1; });
      if (tid < CT) {
        const int k = c0 + tid < nc ? jb.col_idx[c0 + tid] : -1;
        cc[tid] = k < 0 ? 0.f : (GRAD && jb.cc) ? jb.cc[k] : 1.f;
        cvid[tid] = k < 0 ? -2 : k / T2;
      }
      __syncthreads();
      if (!wact) continue;
      const int ntv = min(4, (nc - c0 + 7) >> 3);
      float acc[4][4];
      s_tile(acc, pa_hi, pa_lo, tbase_hi + b_off, tbase_lo + b_off, pitch, Dp, ntv);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h = e >> 1, j = nt * 8 + 2 * t + (e & 1);
          float wgt = rw[h] * cc[j];
          if (jb.excl_same_video && cvid[j] == rvid[h]) wgt = 0.f;
          const float x = wgt != 0.f ? wgt * ex2(acc[nt][e] * c_ex) : 0.f;
          if (GRAD) acc[nt][e] = x;
          else rsum[h] += x;
        }
      }
      if constexpr (GRAD) de_tile<NTD>(o, acc, tbase_hi + v_off, tbase_lo + v_off, pitch, Dp, (ntv + 1) >> 1);
    }
    if constexpr (GRAD) {
#pragma unroll
      for (int n = 0; n < NTD; ++n) {
        const int c = n * 8 + 2 * t;
        if (c < D) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (row[h] >= 0 && (o[n][2 * h] != 0.f || o[n][2 * h + 1] != 0.f))
              red_add2(vec_out + (int64_t)row[h] * D + c, o[n][2 * h] * inv_tau, o[n][2 * h + 1] * inv_tau);
        }
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float s = quad_add(rsum[h]);
        if (t == 0 && row[h] >= 0 && s != 0.f) atomicAdd(sum_out + row[h], s);
      }
    }
  }
}

static size_t meta_bytes(int NC) { return (size_t)(6 * NC + 32) * sizeof(float); }

template <typename K>
static int launch_pair(K kern, size_t& configured, int grid, int threads, size_t smem, int cluster, cudaStream_t st, const PairArgs& A) {
  if (smem > configured) {
    MVF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  if (cluster > 1) launch_kc(kern, grid, threads, smem, st, cluster, A);
  else launch_k(kern, grid, threads, smem, st, A);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaDeviceProp prop;
    sms = (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) ? prop.multiProcessorCount : 148;
  }
  return sms;
}

}  // namespace smma

bool scl_mma_supported(int T, int D) { return T >= 1 && T <= 256 && D >= 4 && D <= 256 && D % 4 == 0; }

// own-pair part of the loss and of its gradient; writes w.c for every row (use_zext: Z += w.zext)
int scl_pair_mma(const float* embs, const int64_t* seq_lens, const int64_t* steps, const float* masks, int Bv, int T, int D,
                 float tau, float two_var, const SclWs& w, int use_zext, float* loss_out, float* d_embs, cudaStream_t st) {
  using namespace smma;
  PairArgs A;
  A.embs = embs; A.seq_lens = seq_lens; A.steps = steps; A.masks = masks; A.T = T; A.D = D; A.tau = tau; A.two_var = two_var;
  A.w = w; A.use_zext = use_zext; A.loss_out = loss_out; A.d_embs = d_embs; A.chunks = 1; A.Wc = 1; A.CS = CT;
  const int Tp = round_up(T, CT), Dp = round_up(D, 32), pitch = Dp * 2 + 16, nbv = (T + 15) / 16;
  const bool wide = Dp > 128;
  const int sms = sm_count();
  const size_t budget = 200 * 1024;
  const size_t tile_b = (size_t)CT * pitch * 2;              // 32 staged rows, hi + lo
  if (Tp == CT) {
    const size_t smem = 2 * tile_b + meta_bytes(2 * Tp);
    static size_t confk[2][4] = {{48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024}, {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024}};
    switch ((T + 7) / 8) {
      case 1: return wide ? launch_pair(scl_pair_mma_kernel<32, true, true, 128, 2, 1>, confk[0][0], Bv, 128, smem, 1, st, A)
                          : launch_pair(scl_pair_mma_kernel<16, true, true, 128, 4, 1>, confk[1][0], Bv, 128, smem, 1, st, A);
      case 2: return wide ? launch_pair(scl_pair_mma_kernel<32, true, true, 128, 2, 2>, confk[0][1], Bv, 128, smem, 1, st, A)
                          : launch_pair(scl_pair_mma_kernel<16, true, true, 128, 4, 2>, confk[1][1], Bv, 128, smem, 1, st, A);
      case 3: return wide ? launch_pair(scl_pair_mma_kernel<32, true, true, 128, 2, 3>, confk[0][2], Bv, 128, smem, 1, st, A)
                          : launch_pair(scl_pair_mma_kernel<16, true, true, 128, 4, 3>, confk[1][2], Bv, 128, smem, 1, st, A);
      default: return wide ? launch_pair(scl_pair_mma_kernel<32, true, true, 128, 2, 4>, confk[0][3], Bv, 128, smem, 1, st, A)
                           : launch_pair(scl_pair_mma_kernel<16, true, true, 128, 4, 4>, confk[1][3], Bv, 128, smem, 1, st, A);
    }
  }
  // one CTA per pair when both views fit and the batch fills the machine
  const size_t both_smem = (size_t)2 * Tp / CT * tile_b + meta_bytes(2 * Tp);
  const int both_warps = 2 * nbv;
  if (both_smem <= budget && both_warps <= (wide ? 8 : 12) && Bv >= sms) {
    static size_t conf[2] = {48 * 1024, 48 * 1024};
    return wide ? launch_pair(scl_pair_mma_kernel<32, true, false, 256, 1>, conf[0], Bv, both_warps * 32, both_smem, 1, st, A)
                : launch_pair(scl_pair_mma_kernel<16, true, false, 384, 1>, conf[1], Bv, both_warps * 32, both_smem, 1, st, A);
  }
  // cluster: 2 views x chunks CTAs; the fewest chunks with <= 8 row blocks per CTA, more when the batch is small
  int chunks = nbv <= 8 ? 1 : 2;
  while (chunks < 4 && Bv * 2 * chunks < sms && (nbv + chunks - 1) / chunks > 1) chunks *= 2;
  const int Wc = (nbv + chunks - 1) / chunks;
  A.chunks = chunks;
  A.Wc = Wc;
  const size_t fixed = (size_t)Wc * 16 * pitch * 2 + meta_bytes(Tp);
  MVF_REQUIRE(fixed + tile_b <= budget, MVF_ERR_UNSUPPORTED, "scl: shared memory for T=%d D=%d", T, D);
  int cst = (int)((budget - fixed) / tile_b);                // partner tiles resident
  if (cst > Tp / CT) cst = Tp / CT;
  A.CS = cst * CT;
  const size_t smem = fixed + (size_t)cst * tile_b;
  int threads = Wc * 32;
  if (Bv * 2 * chunks <= 2 * sms && threads < 256) threads = 256;    // extra warps only help staging
  static size_t conf[2] = {48 * 1024, 48 * 1024};
  return wide ? launch_pair(scl_pair_mma_kernel<32, false, false, 256, 1>, conf[0], Bv * 2 * chunks, threads, smem, 2 * chunks, st, A)
              : launch_pair(scl_pair_mma_kernel<16, false, false, 256, 1>, conf[1], Bv * 2 * chunks, threads, smem, 2 * chunks, st, A);
}

// one launch for up to 4 cross jobs; grad = 0: row sums into sum_out, 1: gradient rows into vec_out
int scl_cross_mma(const float* embs, int N, int T2, int D, float tau, const SclCrossJobs& J, int grad, float* sum_out,
                  float* vec_out, cudaStream_t st) {
  using namespace smma;
  const int Dp = round_up(D, 32), pitch = Dp * 2 + 16;
  const bool wide = Dp > 128;
  const size_t smem = (size_t)(64 + CT) * pitch * 2 + 2 * CT * sizeof(float);
  int gy = cdiv(N, CT);
  if (gy > 64) gy = 64;
  int gx = cdiv(N, 64);
  const int cap = (8 * sm_count() + gy - 1) / gy;
  if (gx > cap) gx = cap;
  const dim3 grid(gx, gy, J.n);
  static size_t conf[4] = {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024};
  auto go = [&](auto kern, size_t& configured) -> int {
    if (smem > configured) {
      MVF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    launch_k(kern, grid, 128, smem, st, embs, D, T2, tau, J, sum_out, vec_out);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  };
  if (grad) return wide ? go(scl_cross_mma_kernel<32, true>, conf[0]) : go(scl_cross_mma_kernel<16, true>, conf[1]);
  return wide ? go(scl_cross_mma_kernel<32, false>, conf[2]) : go(scl_cross_mma_kernel<16, false>, conf[3]);
}

}  // namespace mvf

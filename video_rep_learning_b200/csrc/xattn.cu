// Entity-query cross-attention pooling over the patch tokens of one frame (single head of width SPC):
//   A[e,p] = softmax_p( (Q_s[e]+Q_b) . K[p] / sqrt(SPC) ),  ent[e,:] = sum_p A[e,p] V[p,:]
// (mvformer.py:352-414 + utils.py:11-44; the reference loops over videos, mvformer.py:255-264 -- here one
// CTA per frame, all frames of the batch in one launch).  K|V come from the tcgen05 projection GEMM as one
// [F*P, 2*SPC] matrix.  Both kernels are HBM-bound on the single pass over K|V: rows are read with the
// channel index on the lanes (coalesced), reductions over channels are warp shuffles, reductions over tokens
// run per-thread along the token loop.
#include "kernels.cuh"

namespace mvf {

template <typename T, int EMAX>
__global__ void __launch_bounds__(256)
xattn_fwd_kernel(int P, int E, int SPC, const T* __restrict__ kv, const float* __restrict__ q_s,
                 const float* __restrict__ q_b, float* __restrict__ attn, float* __restrict__ ent, int64_t ld_ent,
                 int one_hot, float p_drop, float inv_keep, DropSeed seed) {
  pdl_entry();
  extern __shared__ float sm[];
  float* Q = sm;                   // [E][SPC]
  float* A = sm + (size_t)E * SPC; // [E][P]
  const int f = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ldkv = 2 * (int64_t)SPC;
  const T* kvf = kv + (int64_t)f * P * ldkv;
  const float scale = rsqrtf((float)SPC);

  for (int i = tid; i < E * SPC; i += blockDim.x) Q[i] = q_s[i] + q_b[i % SPC];
  __syncthreads();

  // scores: one warp per token row
  for (int p = warp; p < P; p += 8) {
    float part[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) part[e] = 0.f;
    const T* kr = kvf + (int64_t)p * ldkv;
    for (int c = lane; c < SPC; c += 32) {
      float k = to_f<T>(kr[c]);
#pragma unroll
      for (int e = 0; e < EMAX; ++e)
        if (e < E) part[e] = fmaf(k, Q[e * SPC + c], part[e]);
    }
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if (e < E) {
        float s = warp_sum(part[e]);
        if (lane == 0) A[e * P + p] = s * scale;
      }
    }
  }
  __syncthreads();

  // softmax over tokens: one warp per entity
  for (int e = warp; e < E; e += 8) {
    float mx = -INFINITY;
    for (int p = lane; p < P; p += 32) mx = fmaxf(mx, A[e * P + p]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int p = lane; p < P; p += 32) {
      float v = expf(A[e * P + p] - mx);
      A[e * P + p] = v;
      sum += v;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int p = lane; p < P; p += 32) {
      float v = A[e * P + p] * inv;
      A[e * P + p] = v;
      if (attn) attn[((int64_t)f * E + e) * P + p] = v;
    }
  }
  __syncthreads();

  // ent[e][c] = sum_p A[e][p] * V[p][c]: one thread per channel
  const int W = SPC + (one_hot ? E : 0);  // logical width of the MLP input (dropout index space)
  for (int c = tid; c < SPC; c += blockDim.x) {
    float acc[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) acc[e] = 0.f;
    const T* vp = kvf + SPC + c;
#pragma unroll 4
    for (int p = 0; p < P; ++p) {
      float v = to_f<T>(vp[(int64_t)p * ldkv]);
#pragma unroll
      for (int e = 0; e < EMAX; ++e)
        if (e < E) acc[e] = fmaf(A[e * P + p], v, acc[e]);
    }
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if (e < E) {
        const int64_t row = (int64_t)f * E + e;
        float v = acc[e];
        if (p_drop > 0.f) v *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + c), p_drop, inv_keep);
        ent[row * ld_ent + c] = v;
      }
    }
  }
  // one-hot entity id columns (mvformer.py:144-149) + zero padding up to ld_ent
  for (int i = tid; i < E * ((int)ld_ent - SPC); i += blockDim.x) {
    const int e = i / ((int)ld_ent - SPC), j = i % ((int)ld_ent - SPC);
    const int64_t row = (int64_t)f * E + e;
    float v = (one_hot && j == e) ? 1.f : 0.f;
    if (v != 0.f && p_drop > 0.f) v *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + SPC + j), p_drop, inv_keep);
    ent[row * ld_ent + SPC + j] = v;
  }
}

template <typename T, int EMAX>
__global__ void __launch_bounds__(256)
xattn_bwd_kernel(int P, int E, int SPC, const T* __restrict__ kv, const float* __restrict__ q_s,
                 const float* __restrict__ q_b, const float* __restrict__ attn, const float* __restrict__ d_ent,
                 int64_t ld_ent, int one_hot, float p_drop, float inv_keep, DropSeed seed, T* __restrict__ d_kv,
                 float* __restrict__ d_q_s, float* __restrict__ d_q_b, float* __restrict__ d_bk,
                 float* __restrict__ d_bv) {
  pdl_entry();
  extern __shared__ float sm[];
  float* Q = sm;                        // [E][SPC]
  float* dEnt = Q + (size_t)E * SPC;    // [E][SPC]
  float* A = dEnt + (size_t)E * SPC;    // [E][P]
  float* dS = A + (size_t)E * P;        // [E][P]  (dA, then dS in place)
  const int f = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ldkv = 2 * (int64_t)SPC;
  const T* kvf = kv + (int64_t)f * P * ldkv;
  T* dkvf = d_kv + (int64_t)f * P * ldkv;
  const float scale = rsqrtf((float)SPC);
  const int W = SPC + (one_hot ? E : 0);

  for (int i = tid; i < E * SPC; i += blockDim.x) {
    const int e = i / SPC, c = i % SPC;
    Q[i] = q_s[i] + q_b[c];
    const int64_t row = (int64_t)f * E + e;
    float g = d_ent[row * ld_ent + c];
    if (p_drop > 0.f) g *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + c), p_drop, inv_keep);
    dEnt[i] = g;
  }
  for (int i = tid; i < E * P; i += blockDim.x) A[i] = attn[(int64_t)f * E * P + i];
  __syncthreads();

  // dA[e][p] = dEnt[e] . V[p]
  for (int p = warp; p < P; p += 8) {
    float part[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) part[e] = 0.f;
    const T* vr = kvf + (int64_t)p * ldkv + SPC;
    for (int c = lane; c < SPC; c += 32) {
      float v = to_f<T>(vr[c]);
#pragma unroll
      for (int e = 0; e < EMAX; ++e)
        if (e < E) part[e] = fmaf(v, dEnt[e * SPC + c], part[e]);
    }
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if (e < E) {
        float s = warp_sum(part[e]);
        if (lane == 0) dS[e * P + p] = s;
      }
    }
  }
  __syncthreads();
  // softmax backward, 1/sqrt(SPC) folded in: dS = A * (dA - <A, dA>) * scale
  for (int e = warp; e < E; e += 8) {
    float dot = 0.f;
    for (int p = lane; p < P; p += 32) dot += A[e * P + p] * dS[e * P + p];
    dot = warp_sum(dot);
    for (int p = lane; p < P; p += 32) dS[e * P + p] = A[e * P + p] * (dS[e * P + p] - dot) * scale;
  }
  __syncthreads();

  // per channel: dK[p][c] = sum_e dS[e][p] Q[e][c];  dV[p][c] = sum_e A[e][p] dEnt[e][c];
  //              dQ[e][c] = sum_p dS[e][p] K[p][c];  bias grads = column sums of dK, dV
  for (int c = tid; c < SPC; c += blockDim.x) {
    float qc[EMAX], gc[EMAX], dq[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      qc[e] = e < E ? Q[e * SPC + c] : 0.f;
      gc[e] = e < E ? dEnt[e * SPC + c] : 0.f;
      dq[e] = 0.f;
    }
    float sbk = 0.f, sbv = 0.f;
#pragma unroll 2
    for (int p = 0; p < P; ++p) {
      const float k = to_f<T>(kvf[(int64_t)p * ldkv + c]);
      float dk = 0.f, dv = 0.f;
#pragma unroll
      for (int e = 0; e < EMAX; ++e) {
        if (e < E) {
          const float ds = dS[e * P + p];
          dk = fmaf(ds, qc[e], dk);
          dv = fmaf(A[e * P + p], gc[e], dv);
          dq[e] = fmaf(ds, k, dq[e]);
        }
      }
      dkvf[(int64_t)p * ldkv + c] = from_f<T>(dk);
      dkvf[(int64_t)p * ldkv + SPC + c] = from_f<T>(dv);
      sbk += dk;
      sbv += dv;
    }
    float dqb = 0.f;
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if (e < E) {
        atomicAdd(d_q_s + e * SPC + c, dq[e]);
        dqb += dq[e];
      }
    }
    atomicAdd(d_q_b + c, dqb);
    atomicAdd(d_bk + c, sbk);
    atomicAdd(d_bv + c, sbv);
  }
}


// =====================================================================================================================
// Single-pass bf16 kernels (E <= 4): every K|V token row (2*SPC bf16, contiguous) is read exactly once with 16-byte
// loads, a warp per row, lanes owning fixed 8-channel chunks (chunk c = lane + 32 j; the first SPC/8 chunks are K, the
// rest V).  Forward runs an online softmax per warp and merges the 8 warp-partials through shared memory; backward
// uses rowdot[e] = <dEnt[e], ent[e]> (= sum_p A dA) from the fp32 pooled output kept by forward, so it needs no
// separate pass to finish the softmax gradient.
// =====================================================================================================================
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

template <int EMAX, int CPL>
__global__ void __launch_bounds__(256, 2)
xattn_fwd_v2_kernel(int P, int E, int SPC, const bf16* __restrict__ kv, const float* __restrict__ q_s,
                    const float* __restrict__ q_b, float* __restrict__ attn, float* __restrict__ ent, int64_t ld_ent,
                    float* __restrict__ ent32, int one_hot, float p_drop, float inv_keep, DropSeed seed) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  float* Q = sm;                         // [E][SPC]
  float* sc = Q + (size_t)E * SPC;       // [E][P] raw scaled scores, later probabilities
  float* wm = sc + (size_t)E * P;        // [8][EMAX]
  float* wl = wm + 8 * EMAX;             // [8][EMAX]
  float* fin = wl + 8 * EMAX;            // [2][EMAX] final max / sum
  float* wacc = fin + 2 * EMAX;          // [8][E][SPC]
  const int f = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KCH = SPC >> 3, NCH = KCH * 2;
  const int64_t ldkv = 2 * (int64_t)SPC;
  const bf16* kvf = kv + (int64_t)f * P * ldkv;
  const float scale = rsqrtf((float)SPC);
  for (int i = tid; i < E * SPC; i += blockDim.x) Q[i] = q_s[i] + q_b[i % SPC];
  __syncthreads();
  // qa[e][j][:] holds the query chunk when chunk (lane + 32 j) is a K chunk, the running A.V accumulator when it is a
  // V chunk (a chunk is never both), which halves the register footprint
  float qa[EMAX][CPL][8];
#pragma unroll
  for (int e = 0; e < EMAX; ++e)
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = lane + 32 * j;
#pragma unroll
      for (int i = 0; i < 8; ++i) qa[e][j][i] = (e < E && c < KCH) ? Q[e * SPC + 8 * c + i] : 0.f;
    }
  float m[EMAX], l[EMAX];
#pragma unroll
  for (int e = 0; e < EMAX; ++e) { m[e] = -INFINITY; l[e] = 0.f; }

  for (int p = warp; p < P; p += 8) {
    const uint4* row = reinterpret_cast<const uint4*>(kvf + (int64_t)p * ldkv);
    float x[CPL][8];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = lane + 32 * j;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (c < NCH) u = __ldg(row + c);
      unpack8(u, x[j]);
    }
    float s[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        if (lane + 32 * j < KCH) {
#pragma unroll
          for (int i = 0; i < 8; ++i) part = fmaf(x[j][i], qa[e][j][i], part);
        }
      }
      s[e] = warp_sum(part) * scale;
    }
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      if (e < E) {
        if (lane == 0) sc[e * P + p] = s[e];
        const float mn = fmaxf(m[e], s[e]);
        const float alpha = expf(m[e] - mn), w = expf(s[e] - mn);
        l[e] = l[e] * alpha + w;
        m[e] = mn;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = lane + 32 * j;
          if (c >= KCH) {
#pragma unroll
            for (int i = 0; i < 8; ++i) qa[e][j][i] = fmaf(w, x[j][i], qa[e][j][i] * alpha);
          }
        }
      }
    }
  }
  // publish warp partials
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
    if (e < E) {
      if (lane == 0) { wm[warp * EMAX + e] = m[e]; wl[warp * EMAX + e] = l[e]; }
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        if (c >= KCH && c < NCH) {
#pragma unroll
          for (int i = 0; i < 8; ++i) wacc[((size_t)warp * E + e) * SPC + 8 * (c - KCH) + i] = qa[e][j][i];
        }
      }
    }
  }
  __syncthreads();
  if (tid < E) {
    float M = -INFINITY;
    for (int w = 0; w < 8; ++w) M = fmaxf(M, wm[w * EMAX + tid]);
    float Ls = 0.f;
    for (int w = 0; w < 8; ++w) Ls += wl[w * EMAX + tid] * expf(wm[w * EMAX + tid] - M);
    fin[tid] = M;
    fin[EMAX + tid] = Ls;
  }
  __syncthreads();
  // probabilities (saved for backward / the attn_holder side channel)
  for (int i = tid; i < E * P; i += blockDim.x) {
    const int e = i / P;
    const float a = expf(sc[i] - fin[e]) / fin[EMAX + e];
    sc[i] = a;
    if (attn) attn[(int64_t)f * E * P + i] = a;
  }
  // pooled entities
  const int W = SPC + (one_hot ? E : 0);
  for (int i = tid; i < E * SPC; i += blockDim.x) {
    const int e = i / SPC, c = i - e * SPC;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += wacc[((size_t)w * E + e) * SPC + c] * expf(wm[w * EMAX + e] - fin[e]);
    v /= fin[EMAX + e];
    const int64_t row = (int64_t)f * E + e;
    if (ent32) ent32[row * SPC + c] = v;
    if (p_drop > 0.f) v *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + c), p_drop, inv_keep);
    ent[row * ld_ent + c] = v;
  }
  for (int i = tid; i < E * ((int)ld_ent - SPC); i += blockDim.x) {
    const int e = i / ((int)ld_ent - SPC), j = i % ((int)ld_ent - SPC);
    const int64_t row = (int64_t)f * E + e;
    float v = (one_hot && j == e) ? 1.f : 0.f;
    if (v != 0.f && p_drop > 0.f) v *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + SPC + j), p_drop, inv_keep);
    ent[row * ld_ent + SPC + j] = v;
  }
}

template <int EMAX, int CPL>
__global__ void __launch_bounds__(256, 1)
xattn_bwd_v2_kernel(int P, int E, int SPC, const bf16* __restrict__ kv, const float* __restrict__ q_s,
                    const float* __restrict__ q_b, const float* __restrict__ attn, const float* __restrict__ d_ent,
                    int64_t ld_ent, const float* __restrict__ ent32, int one_hot, float p_drop, float inv_keep,
                    DropSeed seed, bf16* __restrict__ d_kv, float* __restrict__ d_q_s, float* __restrict__ d_q_b,
                    float* __restrict__ d_bk, float* __restrict__ d_bv) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  float* QG = sm;                          // [E][2*SPC]: Q[e] | dEnt[e]  (same chunk numbering as a K|V row)
  float* A = QG + (size_t)E * 2 * SPC;     // [E][P]
  float* rowdot = A + (size_t)E * P;       // [EMAX]
  float* red = rowdot + EMAX;              // [8][(E+1)][2*SPC] warp partials: dQ (K half) and column sums
  const int f = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KCH = SPC >> 3, NCH = KCH * 2;
  const int64_t ldkv = 2 * (int64_t)SPC;
  const bf16* kvf = kv + (int64_t)f * P * ldkv;
  bf16* dkvf = d_kv + (int64_t)f * P * ldkv;
  const float scale = rsqrtf((float)SPC);
  const int W = SPC + (one_hot ? E : 0);
  for (int i = tid; i < E * SPC; i += blockDim.x) {
    const int e = i / SPC, c = i - e * SPC;
    QG[(size_t)e * 2 * SPC + c] = q_s[i] + q_b[c];
    const int64_t row = (int64_t)f * E + e;
    float g = d_ent[row * ld_ent + c];
    if (p_drop > 0.f) g *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + c), p_drop, inv_keep);
    QG[(size_t)e * 2 * SPC + SPC + c] = g;
  }
  for (int i = tid; i < E * P; i += blockDim.x) A[i] = attn[(int64_t)f * E * P + i];
  __syncthreads();
  for (int e = warp; e < E; e += 8) {
    float d = 0.f;
    for (int c = lane; c < SPC; c += 32) d = fmaf(QG[(size_t)e * 2 * SPC + SPC + c], ent32[((int64_t)f * E + e) * SPC + c], d);
    d = warp_sum(d);
    if (lane == 0) rowdot[e] = d;
  }
  float qg[EMAX][CPL][8], dq[EMAX][CPL][8], bsum[CPL][8];
#pragma unroll
  for (int e = 0; e < EMAX; ++e)
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = lane + 32 * j;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        qg[e][j][i] = (e < E && c < NCH) ? QG[(size_t)e * 2 * SPC + 8 * c + i] : 0.f;
        dq[e][j][i] = 0.f;
      }
    }
#pragma unroll
  for (int j = 0; j < CPL; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[j][i] = 0.f;
  __syncthreads();
  float rd[EMAX];
#pragma unroll
  for (int e = 0; e < EMAX; ++e) rd[e] = e < E ? rowdot[e] : 0.f;

  for (int p = warp; p < P; p += 8) {
    const uint4* row = reinterpret_cast<const uint4*>(kvf + (int64_t)p * ldkv);
    uint4* orow = reinterpret_cast<uint4*>(dkvf + (int64_t)p * ldkv);
    float x[CPL][8];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = lane + 32 * j;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (c < NCH) u = __ldg(row + c);
      unpack8(u, x[j]);
    }
    float ds[EMAX], a[EMAX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e) {
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        if (c >= KCH) {
#pragma unroll
          for (int i = 0; i < 8; ++i) part = fmaf(x[j][i], qg[e][j][i], part);   // V chunk . dEnt chunk
        }
      }
      const float dA = warp_sum(part);
      a[e] = e < E ? A[e * P + p] : 0.f;
      ds[e] = a[e] * (dA - rd[e]) * scale;
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int c = lane + 32 * j;
      if (c < NCH) {
        float o[8];
        if (c < KCH) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float dk = 0.f;
#pragma unroll
            for (int e = 0; e < EMAX; ++e) {
              dk = fmaf(ds[e], qg[e][j][i], dk);
              dq[e][j][i] = fmaf(ds[e], x[j][i], dq[e][j][i]);
            }
            o[i] = dk;
            bsum[j][i] += dk;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float dv = 0.f;
#pragma unroll
            for (int e = 0; e < EMAX; ++e) dv = fmaf(a[e], qg[e][j][i], dv);
            o[i] = dv;
            bsum[j][i] += dv;
          }
        }
        orow[c] = pack8(o);
      }
    }
  }
  // warp partials -> shared -> global atomics (dQ_s, dQ_b, bias gradients)
  const int RW = (E + 1) * 2 * SPC;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    if (c < NCH) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        red[(size_t)warp * RW + (size_t)E * 2 * SPC + 8 * c + i] = bsum[j][i];
        if (c < KCH) {
#pragma unroll
          for (int e = 0; e < EMAX; ++e)
            if (e < E) red[(size_t)warp * RW + (size_t)e * 2 * SPC + 8 * c + i] = dq[e][j][i];
        }
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < 2 * SPC; c += blockDim.x) {
    float b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) b += red[(size_t)w * RW + (size_t)E * 2 * SPC + c];
    if (c < SPC) {
      atomicAdd(d_bk + c, b);
      float qb = 0.f;
      for (int e = 0; e < E; ++e) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[(size_t)w * RW + (size_t)e * 2 * SPC + c];
        atomicAdd(d_q_s + e * SPC + c, v);
        qb += v;
      }
      atomicAdd(d_q_b + c, qb);
    } else {
      atomicAdd(d_bv + (c - SPC), b);
    }
  }
}

static bool v2_ok(int dtype, int E, int SPC, int P, const void* kv, int64_t ld_ent) {
  (void)P; (void)ld_ent;
  return dtype == MVF_BF16 && E <= 4 && SPC % 8 == 0 && (2 * SPC / 8) <= 32 * 4 && ((((uintptr_t)kv) & 15) == 0);
}

template <int EM, int CPL>
static int fwd_v2_launch(int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b, float* attn,
                         void* ent, int64_t ld_ent, float* ent32, int one_hot, float drop_p, DropSeed seed,
                         cudaStream_t st) {
  size_t smem = ((size_t)E * SPC + (size_t)E * P + 16 * EM + 2 * EM + (size_t)8 * E * SPC) * sizeof(float);
  MVF_REQUIRE(smem <= 227 * 1024, MVF_ERR_UNSUPPORTED, "xattn fwd: %zu B of shared memory", smem);
  static bool cfgd = false;
  if (!cfgd) {
    MVF_CHECK_CUDA(cudaFuncSetAttribute(xattn_fwd_v2_kernel<EM, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfgd = true;
  }
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  launch_k(xattn_fwd_v2_kernel<EM, CPL>, F, 256, smem, st, P, E, SPC, (const bf16*)kv, q_s, q_b, attn, (float*)ent, ld_ent, ent32,
                                                     one_hot, drop_p, ik, seed);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
template <int EM, int CPL>
static int bwd_v2_launch(int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b,
                         const float* attn, const void* d_ent, int64_t ld_ent, const float* ent32, int one_hot,
                         float drop_p, DropSeed seed, void* d_kv, float* d_q_s, float* d_q_b, float* d_bk, float* d_bv,
                         cudaStream_t st) {
  size_t smem = ((size_t)E * 2 * SPC + (size_t)E * P + EM + (size_t)8 * (E + 1) * 2 * SPC) * sizeof(float);
  MVF_REQUIRE(smem <= 227 * 1024, MVF_ERR_UNSUPPORTED, "xattn bwd: %zu B of shared memory", smem);
  static bool cfgd = false;
  if (!cfgd) {
    MVF_CHECK_CUDA(cudaFuncSetAttribute(xattn_bwd_v2_kernel<EM, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfgd = true;
  }
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  launch_k(xattn_bwd_v2_kernel<EM, CPL>, F, 256, smem, st, P, E, SPC, (const bf16*)kv, q_s, q_b, attn, (const float*)d_ent, ld_ent,
                                                     ent32, one_hot, drop_p, ik, seed, (bf16*)d_kv, d_q_s, d_q_b, d_bk, d_bv);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

template <typename T>
static int fwd_t(int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b, float* attn, void* ent,
                 int64_t ld_ent, int one_hot, float drop_p, DropSeed seed, cudaStream_t st) {
  size_t smem = ((size_t)E * SPC + (size_t)E * P) * sizeof(float);
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#define LAUNCH_F(EM)                                                                                              \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      MVF_CHECK_CUDA(cudaFuncSetAttribute(xattn_fwd_kernel<T, EM>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)smem));                                                            \
    launch_k(xattn_fwd_kernel<T, EM>, F, 256, smem, st, P, E, SPC, (const T*)kv, q_s, q_b, attn, (float*)ent, ld_ent, one_hot, \
                                                  drop_p, ik, seed);                                              \
  } while (0)
  if (E <= 4) LAUNCH_F(4);
  else if (E <= 8) LAUNCH_F(8);
  else LAUNCH_F(16);
#undef LAUNCH_F
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

int xattn_pool_fwd(int dtype, int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b,
                   float* attn, void* ent, int64_t ld_ent, float* ent32, int one_hot, float drop_p, DropSeed seed,
                   cudaStream_t st) {
  MVF_REQUIRE(E >= 1 && E <= MVF_MAX_ENTITIES, MVF_ERR_UNSUPPORTED, "xattn: %d entities (max %d)", E, MVF_MAX_ENTITIES);
  MVF_REQUIRE(ld_ent >= SPC + (one_hot ? E : 0), MVF_ERR_BAD_ARG, "xattn: ld_ent too small");
  if (F > 0 && ent32 != nullptr && v2_ok(dtype, E, SPC, P, kv, ld_ent)) {
    const int cpl = (2 * SPC / 8 + 31) / 32;
#define MVF_FWD_V2(EM_, CPL_) \
  return fwd_v2_launch<EM_, CPL_>(F, P, E, SPC, kv, q_s, q_b, attn, ent, ld_ent, ent32, one_hot, drop_p, seed, st)
    if (E <= 2) { if (cpl == 1) MVF_FWD_V2(2, 1); if (cpl == 2) MVF_FWD_V2(2, 2); if (cpl == 3) MVF_FWD_V2(2, 3); MVF_FWD_V2(2, 4); }
    if (E == 3) { if (cpl == 1) MVF_FWD_V2(3, 1); if (cpl == 2) MVF_FWD_V2(3, 2); if (cpl == 3) MVF_FWD_V2(3, 3); MVF_FWD_V2(3, 4); }
    if (cpl == 1) MVF_FWD_V2(4, 1); if (cpl == 2) MVF_FWD_V2(4, 2); if (cpl == 3) MVF_FWD_V2(4, 3); MVF_FWD_V2(4, 4);
#undef MVF_FWD_V2
  }
  size_t smem = ((size_t)E * SPC + (size_t)E * P) * sizeof(float);
  MVF_REQUIRE(smem <= 227 * 1024, MVF_ERR_UNSUPPORTED, "xattn fwd: E*(SPC+P) needs %zu B of shared memory", smem);
  if (F <= 0) return MVF_OK;
  if (dtype == MVF_BF16) return fwd_t<bf16>(F, P, E, SPC, kv, q_s, q_b, attn, ent, ld_ent, one_hot, drop_p, seed, st);
  return fwd_t<float>(F, P, E, SPC, kv, q_s, q_b, attn, ent, ld_ent, one_hot, drop_p, seed, st);
}

template <typename T>
static int bwd_t(int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b, const float* attn,
                 const void* d_ent, int64_t ld_ent, int one_hot, float drop_p, DropSeed seed, void* d_kv, float* d_q_s,
                 float* d_q_b, float* d_bk, float* d_bv, cudaStream_t st) {
  size_t smem = (2 * (size_t)E * SPC + 2 * (size_t)E * P) * sizeof(float);
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#define LAUNCH_B(EM)                                                                                              \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      MVF_CHECK_CUDA(cudaFuncSetAttribute(xattn_bwd_kernel<T, EM>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)smem));                                                            \
    launch_k(xattn_bwd_kernel<T, EM>, F, 256, smem, st, P, E, SPC, (const T*)kv, q_s, q_b, attn, (const float*)d_ent, ld_ent, \
                                                  one_hot, drop_p, ik, seed, (T*)d_kv, d_q_s, d_q_b, d_bk, d_bv); \
  } while (0)
  if (E <= 4) LAUNCH_B(4);
  else if (E <= 8) LAUNCH_B(8);
  else LAUNCH_B(16);
#undef LAUNCH_B
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

int xattn_pool_bwd(int dtype, int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b,
                   const float* attn, const void* d_ent, int64_t ld_ent, const float* ent32, int one_hot, float drop_p,
                   DropSeed seed, void* d_kv, float* d_q_s, float* d_q_b, float* d_bk, float* d_bv, cudaStream_t st) {
  MVF_REQUIRE(E >= 1 && E <= MVF_MAX_ENTITIES, MVF_ERR_UNSUPPORTED, "xattn: %d entities (max %d)", E, MVF_MAX_ENTITIES);
  if (F > 0 && ent32 != nullptr && v2_ok(dtype, E, SPC, P, kv, ld_ent) && ((((uintptr_t)d_kv) & 15) == 0)) {
    const int cpl = (2 * SPC / 8 + 31) / 32;
#define MVF_BWD_V2(EM_, CPL_)                                                                                          \
  return bwd_v2_launch<EM_, CPL_>(F, P, E, SPC, kv, q_s, q_b, attn, d_ent, ld_ent, ent32, one_hot, drop_p, seed, d_kv, d_q_s, \
                                  d_q_b, d_bk, d_bv, st)
    if (E <= 2) { if (cpl == 1) MVF_BWD_V2(2, 1); if (cpl == 2) MVF_BWD_V2(2, 2); if (cpl == 3) MVF_BWD_V2(2, 3); MVF_BWD_V2(2, 4); }
    if (E == 3) { if (cpl == 1) MVF_BWD_V2(3, 1); if (cpl == 2) MVF_BWD_V2(3, 2); if (cpl == 3) MVF_BWD_V2(3, 3); MVF_BWD_V2(3, 4); }
    if (cpl == 1) MVF_BWD_V2(4, 1); if (cpl == 2) MVF_BWD_V2(4, 2); if (cpl == 3) MVF_BWD_V2(4, 3); MVF_BWD_V2(4, 4);
#undef MVF_BWD_V2
  }
  size_t smem = (2 * (size_t)E * SPC + 2 * (size_t)E * P) * sizeof(float);
  MVF_REQUIRE(smem <= 227 * 1024, MVF_ERR_UNSUPPORTED, "xattn bwd: E*(SPC+P) needs %zu B of shared memory", smem);
  if (F <= 0) return MVF_OK;
  if (dtype == MVF_BF16)
    return bwd_t<bf16>(F, P, E, SPC, kv, q_s, q_b, attn, d_ent, ld_ent, one_hot, drop_p, seed, d_kv, d_q_s, d_q_b, d_bk,
                       d_bv, st);
  return bwd_t<float>(F, P, E, SPC, kv, q_s, q_b, attn, d_ent, ld_ent, one_hot, drop_p, seed, d_kv, d_q_s, d_q_b, d_bk,
                      d_bv, st);
}

}  // namespace mvf

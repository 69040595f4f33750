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
                 const float* __restrict__ q_b, float* __restrict__ attn, T* __restrict__ ent, int64_t ld_ent, int one_hot,
                 float p_drop, float inv_keep, uint64_t seed) {
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
        ent[row * ld_ent + c] = from_f<T>(v);
      }
    }
  }
  // one-hot entity id columns (mvformer.py:144-149) + zero padding up to ld_ent
  for (int i = tid; i < E * ((int)ld_ent - SPC); i += blockDim.x) {
    const int e = i / ((int)ld_ent - SPC), j = i % ((int)ld_ent - SPC);
    const int64_t row = (int64_t)f * E + e;
    float v = (one_hot && j == e) ? 1.f : 0.f;
    if (v != 0.f && p_drop > 0.f) v *= drop_scale(seed, SITE_FC0, (uint64_t)(row * W + SPC + j), p_drop, inv_keep);
    ent[row * ld_ent + SPC + j] = from_f<T>(v);
  }
}

template <typename T, int EMAX>
__global__ void __launch_bounds__(256)
xattn_bwd_kernel(int P, int E, int SPC, const T* __restrict__ kv, const float* __restrict__ q_s,
                 const float* __restrict__ q_b, const float* __restrict__ attn, const T* __restrict__ d_ent,
                 int64_t ld_ent, int one_hot, float p_drop, float inv_keep, uint64_t seed, T* __restrict__ d_kv,
                 float* __restrict__ d_q_s, float* __restrict__ d_q_b, float* __restrict__ d_bk,
                 float* __restrict__ d_bv) {
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
    float g = to_f<T>(d_ent[row * ld_ent + c]);
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

template <typename T>
static int fwd_t(int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b, float* attn, void* ent,
                 int64_t ld_ent, int one_hot, float drop_p, uint64_t seed, cudaStream_t st) {
  size_t smem = ((size_t)E * SPC + (size_t)E * P) * sizeof(float);
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#define LAUNCH_F(EM)                                                                                              \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      MVF_CHECK_CUDA(cudaFuncSetAttribute(xattn_fwd_kernel<T, EM>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)smem));                                                            \
    xattn_fwd_kernel<T, EM><<<F, 256, smem, st>>>(P, E, SPC, (const T*)kv, q_s, q_b, attn, (T*)ent, ld_ent, one_hot, \
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
                   float* attn, void* ent, int64_t ld_ent, int one_hot, float drop_p, uint64_t seed, cudaStream_t st) {
  MVF_REQUIRE(E >= 1 && E <= MVF_MAX_ENTITIES, MVF_ERR_UNSUPPORTED, "xattn: %d entities (max %d)", E, MVF_MAX_ENTITIES);
  MVF_REQUIRE(ld_ent >= SPC + (one_hot ? E : 0), MVF_ERR_BAD_ARG, "xattn: ld_ent too small");
  size_t smem = ((size_t)E * SPC + (size_t)E * P) * sizeof(float);
  MVF_REQUIRE(smem <= 227 * 1024, MVF_ERR_UNSUPPORTED, "xattn fwd: E*(SPC+P) needs %zu B of shared memory", smem);
  if (F <= 0) return MVF_OK;
  if (dtype == MVF_BF16) return fwd_t<bf16>(F, P, E, SPC, kv, q_s, q_b, attn, ent, ld_ent, one_hot, drop_p, seed, st);
  return fwd_t<float>(F, P, E, SPC, kv, q_s, q_b, attn, ent, ld_ent, one_hot, drop_p, seed, st);
}

template <typename T>
static int bwd_t(int F, int P, int E, int SPC, const void* kv, const float* q_s, const float* q_b, const float* attn,
                 const void* d_ent, int64_t ld_ent, int one_hot, float drop_p, uint64_t seed, void* d_kv, float* d_q_s,
                 float* d_q_b, float* d_bk, float* d_bv, cudaStream_t st) {
  size_t smem = (2 * (size_t)E * SPC + 2 * (size_t)E * P) * sizeof(float);
  float ik = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#define LAUNCH_B(EM)                                                                                              \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      MVF_CHECK_CUDA(cudaFuncSetAttribute(xattn_bwd_kernel<T, EM>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)smem));                                                            \
    xattn_bwd_kernel<T, EM><<<F, 256, smem, st>>>(P, E, SPC, (const T*)kv, q_s, q_b, attn, (const T*)d_ent, ld_ent, \
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
                   const float* attn, const void* d_ent, int64_t ld_ent, int one_hot, float drop_p, uint64_t seed,
                   void* d_kv, float* d_q_s, float* d_q_b, float* d_bk, float* d_bv, cudaStream_t st) {
  MVF_REQUIRE(E >= 1 && E <= MVF_MAX_ENTITIES, MVF_ERR_UNSUPPORTED, "xattn: %d entities (max %d)", E, MVF_MAX_ENTITIES);
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

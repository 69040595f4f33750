// Temporal multi-head self-attention core (models/utils.py:11-44, 87-103): softmax(Q K^T / sqrt(dk) + keymask) V
// per (view, head), streaming over key tiles with an online softmax so the [S,S] score matrix never reaches
// HBM (the reference materialises [BV,8,S,S]).  fp32 math on CUDA cores: with dk = 32 the exp, not the MMA, is
// the bound.  These are the exact-fp32 kernels (fp32 tokens / parity mode); on the tensor-core
// backend S <= 64 runs in attention_tc.cu (mma.sync) and longer sequences in attention_fa.cu (tcgen05 / TMEM / TMA).
//
// Layout: qkv [B*S, 3*H] rows = (view, token), columns Q | K | V with head h at [h*dk, (h+1)*dk).
// One thread owns one query row (forward, dQ) or one key row (dK/dV); key/query tiles are staged in shared
// memory and read as warp-wide broadcasts.
#include <stdlib.h>

#include "kernels.cuh"

namespace mvf {

constexpr int ATT_THREADS = 64;  // rows per CTA
constexpr int ATT_TILE = 32;     // keys (or queries) per shared-memory tile

template <typename T, int DK>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(int S, int H, const T* __restrict__ qkv, const float* __restrict__ keymask, T* __restrict__ ctx,
                float* __restrict__ lse) {
  pdl_entry();
  __shared__ __align__(16) float Ks[ATT_TILE][DK];
  __shared__ __align__(16) float Vs[ATT_TILE][DK];
  __shared__ float Ms[ATT_TILE];
  const int b = blockIdx.z, h = blockIdx.y, heads = gridDim.y;
  const int i = blockIdx.x * ATT_THREADS + threadIdx.x;
  const bool active = i < S;
  const int64_t ld = 3 * (int64_t)H;
  const T* base = qkv + (int64_t)b * S * ld + h * DK;
  const float scale = 1.0f / sqrtf((float)DK);

  float q[DK], o[DK];
#pragma unroll
  for (int d = 0; d < DK; ++d) {
    q[d] = active ? to_f<T>(base[(int64_t)i * ld + d]) : 0.f;
    o[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;

  for (int j0 = 0; j0 < S; j0 += ATT_TILE) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < ATT_TILE * DK; idx += ATT_THREADS) {
      const int jj = idx / DK, d = idx % DK;
      const int j = j0 + jj;
      float kvv = 0.f, vvv = 0.f;
      if (j < S) {
        kvv = to_f<T>(base[(int64_t)j * ld + H + d]);
        vvv = to_f<T>(base[(int64_t)j * ld + 2 * H + d]);
      }
      Ks[jj][d] = kvv;
      Vs[jj][d] = vvv;
    }
    if (threadIdx.x < ATT_TILE) {
      const int j = j0 + threadIdx.x;
      Ms[threadIdx.x] = (j < S && (keymask == nullptr || keymask[(int64_t)b * S + j] != 0.f)) ? 1.f : 0.f;
    }
    __syncthreads();
    float s[ATT_TILE];
    float tmax = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < ATT_TILE; ++jj) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // four independent chains hide the FMA latency
#pragma unroll
      for (int d = 0; d < DK; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[jj][d]);
        a0 = fmaf(q[d], kk.x, a0);
        a1 = fmaf(q[d + 1], kk.y, a1);
        a2 = fmaf(q[d + 2], kk.z, a2);
        a3 = fmaf(q[d + 3], kk.w, a3);
      }
      float acc = (a0 + a1) + (a2 + a3);
      acc = Ms[jj] != 0.f ? acc * scale : -INFINITY;
      s[jj] = acc;
      tmax = fmaxf(tmax, acc);
    }
    const float m_new = fmaxf(m, tmax);
    if (m_new == -INFINITY) continue;  // every key so far is masked
    const float alpha = (m == -INFINITY) ? 0.f : expf(m - m_new);
    l *= alpha;
#pragma unroll
    for (int d = 0; d < DK; ++d) o[d] *= alpha;
#pragma unroll
    for (int jj = 0; jj < ATT_TILE; ++jj) {
      const float p = expf(s[jj] - m_new);  // exp(-inf) = 0 for masked keys
      l += p;
#pragma unroll
      for (int d = 0; d < DK; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[jj][d]);
        o[d] = fmaf(p, vv.x, o[d]);
        o[d + 1] = fmaf(p, vv.y, o[d + 1]);
        o[d + 2] = fmaf(p, vv.z, o[d + 2]);
        o[d + 3] = fmaf(p, vv.w, o[d + 3]);
      }
    }
    m = m_new;
  }
  if (active) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    T* out = ctx + ((int64_t)b * S + i) * H + h * DK;
#pragma unroll
    for (int d = 0; d < DK; ++d) out[d] = from_f<T>(o[d] * inv);
    lse[((int64_t)b * heads + h) * S + i] = m + logf(l);
  }
}

// dQ (and delta = rowsum(dO * O), written for the dK/dV kernel)
template <typename T, int DK>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(int S, int H, const T* __restrict__ qkv, const float* __restrict__ keymask,
                   const T* __restrict__ ctx, const float* __restrict__ lse, const T* __restrict__ d_ctx,
                   T* __restrict__ d_qkv, float* __restrict__ delta) {
  pdl_entry();
  __shared__ __align__(16) float Ks[ATT_TILE][DK];
  __shared__ __align__(16) float Vs[ATT_TILE][DK];
  __shared__ float Ms[ATT_TILE];
  const int b = blockIdx.z, h = blockIdx.y, heads = gridDim.y;
  const int i = blockIdx.x * ATT_THREADS + threadIdx.x;
  const bool active = i < S;
  const int64_t ld = 3 * (int64_t)H;
  const T* base = qkv + (int64_t)b * S * ld + h * DK;
  const float scale = 1.0f / sqrtf((float)DK);

  float q[DK], go[DK], dq[DK];
  float dl = 0.f;
#pragma unroll
  for (int d = 0; d < DK; ++d) {
    q[d] = active ? to_f<T>(base[(int64_t)i * ld + d]) : 0.f;
    go[d] = active ? to_f<T>(d_ctx[((int64_t)b * S + i) * H + h * DK + d]) : 0.f;
    float ov = active ? to_f<T>(ctx[((int64_t)b * S + i) * H + h * DK + d]) : 0.f;
    dl = fmaf(go[d], ov, dl);
    dq[d] = 0.f;
  }
  const float lse_i = active ? lse[((int64_t)b * heads + h) * S + i] : 0.f;
  if (active) delta[((int64_t)b * heads + h) * S + i] = dl;

  for (int j0 = 0; j0 < S; j0 += ATT_TILE) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < ATT_TILE * DK; idx += ATT_THREADS) {
      const int jj = idx / DK, d = idx % DK;
      const int j = j0 + jj;
      float kvv = 0.f, vvv = 0.f;
      if (j < S) {
        kvv = to_f<T>(base[(int64_t)j * ld + H + d]);
        vvv = to_f<T>(base[(int64_t)j * ld + 2 * H + d]);
      }
      Ks[jj][d] = kvv;
      Vs[jj][d] = vvv;
    }
    if (threadIdx.x < ATT_TILE) {
      const int j = j0 + threadIdx.x;
      Ms[threadIdx.x] = (j < S && (keymask == nullptr || keymask[(int64_t)b * S + j] != 0.f)) ? 1.f : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < ATT_TILE; ++jj) {
      if (Ms[jj] == 0.f) continue;  // block-uniform
      float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;
#pragma unroll
      for (int d = 0; d < DK; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[jj][d]);
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[jj][d]);
        s0 = fmaf(q[d], kk.x, s0);
        s1 = fmaf(q[d + 1], kk.y, s1);
        s0 = fmaf(q[d + 2], kk.z, s0);
        s1 = fmaf(q[d + 3], kk.w, s1);
        p0 = fmaf(go[d], vv.x, p0);
        p1 = fmaf(go[d + 1], vv.y, p1);
        p0 = fmaf(go[d + 2], vv.z, p0);
        p1 = fmaf(go[d + 3], vv.w, p1);
      }
      const float p = expf((s0 + s1) * scale - lse_i);
      const float ds = p * ((p0 + p1) - dl) * scale;
#pragma unroll
      for (int d = 0; d < DK; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[jj][d]);
        dq[d] = fmaf(ds, kk.x, dq[d]);
        dq[d + 1] = fmaf(ds, kk.y, dq[d + 1]);
        dq[d + 2] = fmaf(ds, kk.z, dq[d + 2]);
        dq[d + 3] = fmaf(ds, kk.w, dq[d + 3]);
      }
    }
  }
  if (active) {
    T* out = d_qkv + ((int64_t)b * S + i) * ld + h * DK;
#pragma unroll
    for (int d = 0; d < DK; ++d) out[d] = from_f<T>(dq[d]);
  }
}

// dK, dV: one thread per key row, query tiles broadcast from shared memory
template <typename T, int DK>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(int S, int H, const T* __restrict__ qkv, const float* __restrict__ keymask,
                    const float* __restrict__ lse, const T* __restrict__ d_ctx, const float* __restrict__ delta,
                    T* __restrict__ d_qkv) {
  pdl_entry();
  __shared__ __align__(16) float Qs[ATT_TILE][DK];
  __shared__ __align__(16) float Gs[ATT_TILE][DK];
  __shared__ float Ls[ATT_TILE];
  __shared__ float Ds[ATT_TILE];
  const int b = blockIdx.z, h = blockIdx.y, heads = gridDim.y;
  const int j = blockIdx.x * ATT_THREADS + threadIdx.x;
  const int64_t ld = 3 * (int64_t)H;
  const T* base = qkv + (int64_t)b * S * ld + h * DK;
  const float scale = 1.0f / sqrtf((float)DK);
  const bool active = j < S && (keymask == nullptr || keymask[(int64_t)b * S + j] != 0.f);

  float k[DK], v[DK], dk[DK], dv[DK];
#pragma unroll
  for (int d = 0; d < DK; ++d) {
    k[d] = active ? to_f<T>(base[(int64_t)j * ld + H + d]) : 0.f;
    v[d] = active ? to_f<T>(base[(int64_t)j * ld + 2 * H + d]) : 0.f;
    dk[d] = 0.f;
    dv[d] = 0.f;
  }
  for (int i0 = 0; i0 < S; i0 += ATT_TILE) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < ATT_TILE * DK; idx += ATT_THREADS) {
      const int ii = idx / DK, d = idx % DK;
      const int i = i0 + ii;
      float qv = 0.f, gv = 0.f;
      if (i < S) {
        qv = to_f<T>(base[(int64_t)i * ld + d]);
        gv = to_f<T>(d_ctx[((int64_t)b * S + i) * H + h * DK + d]);
      }
      Qs[ii][d] = qv;
      Gs[ii][d] = gv;
    }
    if (threadIdx.x < ATT_TILE) {
      const int i = i0 + threadIdx.x;
      Ls[threadIdx.x] = i < S ? lse[((int64_t)b * heads + h) * S + i] : INFINITY;  // exp(s - inf) = 0
      Ds[threadIdx.x] = i < S ? delta[((int64_t)b * heads + h) * S + i] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
#pragma unroll 4
    for (int ii = 0; ii < ATT_TILE; ++ii) {
      float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;
#pragma unroll
      for (int d = 0; d < DK; d += 4) {
        const float4 qq = *reinterpret_cast<const float4*>(&Qs[ii][d]);
        const float4 gg = *reinterpret_cast<const float4*>(&Gs[ii][d]);
        s0 = fmaf(qq.x, k[d], s0);
        s1 = fmaf(qq.y, k[d + 1], s1);
        s0 = fmaf(qq.z, k[d + 2], s0);
        s1 = fmaf(qq.w, k[d + 3], s1);
        p0 = fmaf(gg.x, v[d], p0);
        p1 = fmaf(gg.y, v[d + 1], p1);
        p0 = fmaf(gg.z, v[d + 2], p0);
        p1 = fmaf(gg.w, v[d + 3], p1);
      }
      const float p = expf((s0 + s1) * scale - Ls[ii]);
      const float ds = p * ((p0 + p1) - Ds[ii]) * scale;
#pragma unroll
      for (int d = 0; d < DK; d += 4) {
        const float4 qq = *reinterpret_cast<const float4*>(&Qs[ii][d]);
        const float4 gg = *reinterpret_cast<const float4*>(&Gs[ii][d]);
        dv[d] = fmaf(p, gg.x, dv[d]);
        dv[d + 1] = fmaf(p, gg.y, dv[d + 1]);
        dv[d + 2] = fmaf(p, gg.z, dv[d + 2]);
        dv[d + 3] = fmaf(p, gg.w, dv[d + 3]);
        dk[d] = fmaf(ds, qq.x, dk[d]);
        dk[d + 1] = fmaf(ds, qq.y, dk[d + 1]);
        dk[d + 2] = fmaf(ds, qq.z, dk[d + 2]);
        dk[d + 3] = fmaf(ds, qq.w, dk[d + 3]);
      }
    }
  }
  if (j < S) {
    T* out = d_qkv + ((int64_t)b * S + j) * ld + h * DK;
#pragma unroll
    for (int d = 0; d < DK; ++d) {
      out[H + d] = from_f<T>(dk[d]);
      out[2 * H + d] = from_f<T>(dv[d]);
    }
  }
}

// =====================================================================================================================
// Short sequences (S <= 128, the temporal encoder at S = E*T = 60 .. 128): one CTA per (view, head) with K, V (and Q, dO
// in backward) resident in shared memory and FOUR threads per row, each owning DK/4 channels (dot products finished with
// two quad shuffles).  The row-per-thread kernels above leave the machine at ~7 warps per SM at these sizes (measured
// 25 us forward, 79 us backward per layer at B = 64); four threads per row quadruple the parallelism, and the fused
// backward computes dQ, dK and dV in one launch.
// =====================================================================================================================
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

template <int DK>
__global__ void __launch_bounds__(512)
attn_fwd_quad_kernel(int S, int H, const float* __restrict__ qkv, const float* __restrict__ keymask,
                     float* __restrict__ ctx, float* __restrict__ lse) {
  pdl_entry();
  constexpr int DQ = DK / 4, D4 = DK / 4;
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;
  float* Vs = Ks + S * DK;
  float* Ms = Vs + S * DK;
  const int h = blockIdx.x, b = blockIdx.y, heads = gridDim.x, tid = threadIdx.x;
  const int64_t ld = 3 * (int64_t)H;
  const float* base = qkv + (int64_t)b * S * ld + h * DK;
  for (int idx = tid; idx < S * D4; idx += blockDim.x) {
    const int j = idx / D4, d4 = idx - j * D4;
    *reinterpret_cast<float4*>(Ks + j * DK + 4 * d4) = *reinterpret_cast<const float4*>(base + (int64_t)j * ld + H + 4 * d4);
    *reinterpret_cast<float4*>(Vs + j * DK + 4 * d4) = *reinterpret_cast<const float4*>(base + (int64_t)j * ld + 2 * H + 4 * d4);
  }
  for (int j = tid; j < S; j += blockDim.x) Ms[j] = (keymask == nullptr || keymask[(int64_t)b * S + j] != 0.f) ? 1.f : 0.f;
  __syncthreads();
  const int i = tid >> 2, c = tid & 3;
  const bool active = i < S;
  // scores are kept in the log2 domain (scale * log2 e folded into q): exp2f is a single MUFU instruction + fix-ups,
  // expf is ~4x the instructions, and every thread of a quad evaluates the row's softmax terms
  const float scale = 1.4426950408889634f / sqrtf((float)DK);
  float q[DQ], o[DQ];
#pragma unroll
  for (int d = 0; d < DQ; d += 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) v = *reinterpret_cast<const float4*>(base + (int64_t)i * ld + c * DQ + d);
    q[d] = v.x; q[d + 1] = v.y; q[d + 2] = v.z; q[d + 3] = v.w;
    o[d] = o[d + 1] = o[d + 2] = o[d + 3] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < S; j0 += 4) {
    float sc[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = min(j0 + jj, S - 1);
      const float* kr = Ks + j * DK + c * DQ;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int d = 0; d < DQ; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + d);
        a0 = fmaf(q[d], kk.x, a0); a1 = fmaf(q[d + 1], kk.y, a1);
        a0 = fmaf(q[d + 2], kk.z, a0); a1 = fmaf(q[d + 3], kk.w, a1);
      }
      sc[jj] = a0 + a1;
    }
    float tmax = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float t = quad_sum(sc[jj]);
      const int j = j0 + jj;
      sc[jj] = (j < S && Ms[min(j, S - 1)] != 0.f) ? t * scale : -INFINITY;
      tmax = fmaxf(tmax, sc[jj]);
    }
    const float m_new = fmaxf(m, tmax);
    if (m_new == -INFINITY) continue;   // every key so far is masked (uniform within the quad)
    const float alpha = (m == -INFINITY) ? 0.f : exp2f(m - m_new);
    l *= alpha;
#pragma unroll
    for (int d = 0; d < DQ; ++d) o[d] *= alpha;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float p = exp2f(sc[jj] - m_new);   // 2^-inf = 0 for masked / out-of-range keys
      l += p;
      const float* vr = Vs + min(j0 + jj, S - 1) * DK + c * DQ;
#pragma unroll
      for (int d = 0; d < DQ; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(vr + d);
        o[d] = fmaf(p, vv.x, o[d]); o[d + 1] = fmaf(p, vv.y, o[d + 1]);
        o[d + 2] = fmaf(p, vv.z, o[d + 2]); o[d + 3] = fmaf(p, vv.w, o[d + 3]);
      }
    }
    m = m_new;
  }
  if (active) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    float* out = ctx + ((int64_t)b * S + i) * H + h * DK + c * DQ;
#pragma unroll
    for (int d = 0; d < DQ; d += 4)
      *reinterpret_cast<float4*>(out + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
    if (c == 0) lse[((int64_t)b * heads + h) * S + i] = m * 0.6931471805599453f + logf(l);   // natural-log units
  }
}

template <int DK>
__global__ void __launch_bounds__(512)
attn_bwd_quad_kernel(int S, int H, const float* __restrict__ qkv, const float* __restrict__ keymask,
                     const float* __restrict__ ctx, const float* __restrict__ lse, const float* __restrict__ d_ctx,
                     float* __restrict__ d_qkv) {
  pdl_entry();
  constexpr int DQ = DK / 4, D4 = DK / 4;
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;
  float* Ks = Qs + S * DK;
  float* Vs = Ks + S * DK;
  float* Gs = Vs + S * DK;
  float* Ls = Gs + S * DK;
  float* Ds = Ls + S;
  float* Ms = Ds + S;
  const int h = blockIdx.x, b = blockIdx.y, heads = gridDim.x, tid = threadIdx.x;
  const int64_t ld = 3 * (int64_t)H;
  const float* base = qkv + (int64_t)b * S * ld + h * DK;
  const float* gbase = d_ctx + (int64_t)b * S * H + h * DK;
  for (int idx = tid; idx < S * D4; idx += blockDim.x) {
    const int j = idx / D4, d4 = idx - j * D4;
    *reinterpret_cast<float4*>(Qs + j * DK + 4 * d4) = *reinterpret_cast<const float4*>(base + (int64_t)j * ld + 4 * d4);
    *reinterpret_cast<float4*>(Ks + j * DK + 4 * d4) = *reinterpret_cast<const float4*>(base + (int64_t)j * ld + H + 4 * d4);
    *reinterpret_cast<float4*>(Vs + j * DK + 4 * d4) = *reinterpret_cast<const float4*>(base + (int64_t)j * ld + 2 * H + 4 * d4);
    *reinterpret_cast<float4*>(Gs + j * DK + 4 * d4) = *reinterpret_cast<const float4*>(gbase + (int64_t)j * H + 4 * d4);
  }
  for (int j = tid; j < S; j += blockDim.x) {
    Ms[j] = (keymask == nullptr || keymask[(int64_t)b * S + j] != 0.f) ? 1.f : 0.f;
    Ls[j] = lse[((int64_t)b * heads + h) * S + j] * 1.4426950408889634f;
  }
  __syncthreads();
  const int r = tid >> 2, c = tid & 3;      // row (query in phase A, key in phase B) and channel quarter
  const bool inrange = r < S;
  const int rr = min(r, S - 1);
  const float scale = 1.0f / sqrtf((float)DK);
  const float scale2 = scale * 1.4426950408889634f;   // log2 domain for exp2f; Ls is converted once below
  // delta_i = <dO_i, O_i>
  {
    float dl = 0.f;
    if (inrange) {
      const float* orow = ctx + ((int64_t)b * S + r) * H + h * DK + c * DQ;
#pragma unroll
      for (int d = 0; d < DQ; d += 4) {
        const float4 ov = *reinterpret_cast<const float4*>(orow + d);
        const float4 gv = *reinterpret_cast<const float4*>(Gs + r * DK + c * DQ + d);
        dl = fmaf(gv.x, ov.x, dl); dl = fmaf(gv.y, ov.y, dl); dl = fmaf(gv.z, ov.z, dl); dl = fmaf(gv.w, ov.w, dl);
      }
    }
    dl = quad_sum(dl);
    if (inrange && c == 0) Ds[r] = dl;
  }
  __syncthreads();
  float* orow_out = d_qkv + ((int64_t)b * S + rr) * ld + h * DK + c * DQ;
  // ---- phase A: dQ_i = sum_j dS_ij K_j ----
  {
    float q[DQ], g[DQ], dq[DQ];
#pragma unroll
    for (int d = 0; d < DQ; ++d) { q[d] = Qs[rr * DK + c * DQ + d]; g[d] = Gs[rr * DK + c * DQ + d]; dq[d] = 0.f; }
    const float li = Ls[rr], di = Ds[rr];
    for (int j = 0; j < S; ++j) {
      if (Ms[j] == 0.f) continue;   // block-uniform
      const float* kr = Ks + j * DK + c * DQ;
      const float* vr = Vs + j * DK + c * DQ;
      float s0 = 0.f, p0 = 0.f;
#pragma unroll
      for (int d = 0; d < DQ; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + d);
        const float4 vv = *reinterpret_cast<const float4*>(vr + d);
        s0 = fmaf(q[d], kk.x, s0); s0 = fmaf(q[d + 1], kk.y, s0); s0 = fmaf(q[d + 2], kk.z, s0); s0 = fmaf(q[d + 3], kk.w, s0);
        p0 = fmaf(g[d], vv.x, p0); p0 = fmaf(g[d + 1], vv.y, p0); p0 = fmaf(g[d + 2], vv.z, p0); p0 = fmaf(g[d + 3], vv.w, p0);
      }
      s0 = quad_sum(s0);
      p0 = quad_sum(p0);
      const float p = exp2f(s0 * scale2 - li);
      const float ds = p * (p0 - di) * scale;
#pragma unroll
      for (int d = 0; d < DQ; d += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(kr + d);
        dq[d] = fmaf(ds, kk.x, dq[d]); dq[d + 1] = fmaf(ds, kk.y, dq[d + 1]);
        dq[d + 2] = fmaf(ds, kk.z, dq[d + 2]); dq[d + 3] = fmaf(ds, kk.w, dq[d + 3]);
      }
    }
    if (inrange) {
#pragma unroll
      for (int d = 0; d < DQ; d += 4) *reinterpret_cast<float4*>(orow_out + d) = make_float4(dq[d], dq[d + 1], dq[d + 2], dq[d + 3]);
    }
  }
  // ---- phase B: dK_j = sum_i dS_ij Q_i, dV_j = sum_i P_ij dO_i ----
  {
    const bool kactive = inrange && Ms[rr] != 0.f;
    float k[DQ], v[DQ], dk[DQ], dv[DQ];
#pragma unroll
    for (int d = 0; d < DQ; ++d) {
      k[d] = kactive ? Ks[rr * DK + c * DQ + d] : 0.f;
      v[d] = kactive ? Vs[rr * DK + c * DQ + d] : 0.f;
      dk[d] = 0.f; dv[d] = 0.f;
    }
    for (int i = 0; i < S; ++i) {
      const float* qr = Qs + i * DK + c * DQ;
      const float* gr = Gs + i * DK + c * DQ;
      float s0 = 0.f, p0 = 0.f;
#pragma unroll
      for (int d = 0; d < DQ; d += 4) {
        const float4 qq = *reinterpret_cast<const float4*>(qr + d);
        const float4 gg = *reinterpret_cast<const float4*>(gr + d);
        s0 = fmaf(qq.x, k[d], s0); s0 = fmaf(qq.y, k[d + 1], s0); s0 = fmaf(qq.z, k[d + 2], s0); s0 = fmaf(qq.w, k[d + 3], s0);
        p0 = fmaf(gg.x, v[d], p0); p0 = fmaf(gg.y, v[d + 1], p0); p0 = fmaf(gg.z, v[d + 2], p0); p0 = fmaf(gg.w, v[d + 3], p0);
      }
      s0 = quad_sum(s0);
      p0 = quad_sum(p0);
      const float p = kactive ? exp2f(s0 * scale2 - Ls[i]) : 0.f;
      const float ds = p * (p0 - Ds[i]) * scale;
#pragma unroll
      for (int d = 0; d < DQ; d += 4) {
        const float4 qq = *reinterpret_cast<const float4*>(qr + d);
        const float4 gg = *reinterpret_cast<const float4*>(gr + d);
        dv[d] = fmaf(p, gg.x, dv[d]); dv[d + 1] = fmaf(p, gg.y, dv[d + 1]); dv[d + 2] = fmaf(p, gg.z, dv[d + 2]); dv[d + 3] = fmaf(p, gg.w, dv[d + 3]);
        dk[d] = fmaf(ds, qq.x, dk[d]); dk[d + 1] = fmaf(ds, qq.y, dk[d + 1]); dk[d + 2] = fmaf(ds, qq.z, dk[d + 2]); dk[d + 3] = fmaf(ds, qq.w, dk[d + 3]);
      }
    }
    if (inrange) {
#pragma unroll
      for (int d = 0; d < DQ; d += 4) {
        *reinterpret_cast<float4*>(orow_out + H + d) = make_float4(dk[d], dk[d + 1], dk[d + 2], dk[d + 3]);
        *reinterpret_cast<float4*>(orow_out + 2 * H + d) = make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]);
      }
    }
  }
}

static bool quad_ok(int dtype, int S, int dk, const void* a, const void* b2) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MVF_ATTN_QUAD");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on && dtype == MVF_F32 && S <= 128 && (dk == 32 || dk == 64) && ((((uintptr_t)a) & 15) == 0) && ((((uintptr_t)b2) & 15) == 0);
}
template <int DK>
static int fwd_quad_launch(int B, int S, int heads, const void* qkv, const float* keymask, void* ctx, float* lse, cudaStream_t st) {
  const size_t smem = ((size_t)2 * S * DK + S) * sizeof(float);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    MVF_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_quad_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int nt = (int)round_up(4 * S, 32);
  launch_k(attn_fwd_quad_kernel<DK>, dim3(heads, B), nt, smem, st, S, heads * DK, (const float*)qkv, keymask, (float*)ctx, lse);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
template <int DK>
static int bwd_quad_launch(int B, int S, int heads, const void* qkv, const float* keymask, const void* ctx, const float* lse,
                           const void* d_ctx, void* d_qkv, cudaStream_t st) {
  const size_t smem = ((size_t)4 * S * DK + 3 * S) * sizeof(float);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    MVF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_quad_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int nt = (int)round_up(4 * S, 32);
  launch_k(attn_bwd_quad_kernel<DK>, dim3(heads, B), nt, smem, st, S, heads * DK, (const float*)qkv, keymask, (const float*)ctx, lse,
                                                             (const float*)d_ctx, (float*)d_qkv);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

template <typename T, int DK>
static int fwd_launch(int B, int S, int heads, const void* qkv, const float* keymask, void* ctx, float* lse,
                      cudaStream_t st) {
  dim3 grid(cdiv(S, ATT_THREADS), heads, B);
  launch_k(attn_fwd_kernel<T, DK>, grid, ATT_THREADS, 0, st, S, heads * DK, (const T*)qkv, keymask, (T*)ctx, lse);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
template <typename T, int DK>
static int bwd_launch(int B, int S, int heads, const void* qkv, const float* keymask, const void* ctx, const float* lse,
                      const void* d_ctx, void* d_qkv, float* delta, cudaStream_t st) {
  dim3 grid(cdiv(S, ATT_THREADS), heads, B);
  launch_k(attn_bwd_dq_kernel<T, DK>, grid, ATT_THREADS, 0, st, S, heads * DK, (const T*)qkv, keymask, (const T*)ctx, lse,
                                                          (const T*)d_ctx, (T*)d_qkv, delta);
  MVF_CHECK_LAUNCH();
  launch_k(attn_bwd_dkv_kernel<T, DK>, grid, ATT_THREADS, 0, st, S, heads * DK, (const T*)qkv, keymask, lse, (const T*)d_ctx,
                                                           delta, (T*)d_qkv);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

#define DISPATCH_DK(FN, ...)                                                                   \
  do {                                                                                         \
    if (dtype == MVF_BF16) {                                                                   \
      if (dk == 8) return FN<bf16, 8>(__VA_ARGS__);                                            \
      if (dk == 16) return FN<bf16, 16>(__VA_ARGS__);                                          \
      if (dk == 32) return FN<bf16, 32>(__VA_ARGS__);                                          \
      if (dk == 64) return FN<bf16, 64>(__VA_ARGS__);                                          \
    } else {                                                                                   \
      if (dk == 8) return FN<float, 8>(__VA_ARGS__);                                           \
      if (dk == 16) return FN<float, 16>(__VA_ARGS__);                                         \
      if (dk == 32) return FN<float, 32>(__VA_ARGS__);                                         \
      if (dk == 64) return FN<float, 64>(__VA_ARGS__);                                         \
    }                                                                                          \
    set_error("attention: head width %d not supported (8, 16, 32, 64)", dk);                  \
    return MVF_ERR_UNSUPPORTED;                                                                \
  } while (0)

size_t attention_ws_bytes(int B, int S, int heads, int dk) {
  return (dk == 32 && tc_available()) ? attention_fa_ws_bytes(B, S, heads) : 0;
}
int attention_fwd(int dtype, int B, int S, int heads, int dk, const void* qkv, const float* keymask, void* ctx,
                  float* lse, cudaStream_t st, bool allow_split, void* ws, size_t ws_bytes) {
  if (B <= 0 || S <= 0) return MVF_OK;
  MVF_REQUIRE(B <= 65535 && heads <= 65535, MVF_ERR_BAD_ARG, "attention: grid too large");
  if (attention_fa_ok(dtype, S, dk, heads * dk, allow_split, ws, ws_bytes, B, heads) && ((((uintptr_t)qkv) | ((uintptr_t)ctx)) & 15) == 0)
    return attention_fa_fwd(B, S, heads, qkv, keymask, ctx, lse, ws, st);
  if (attention_tc_ok(dtype, S, dk, heads * dk, qkv, ctx, allow_split)) return attention_tc_fwd(B, S, heads, qkv, keymask, ctx, lse, st);
  if (quad_ok(dtype, S, dk, qkv, ctx) && (heads * dk) % 4 == 0) {
    if (dk == 32) return fwd_quad_launch<32>(B, S, heads, qkv, keymask, ctx, lse, st);
    return fwd_quad_launch<64>(B, S, heads, qkv, keymask, ctx, lse, st);
  }
  DISPATCH_DK(fwd_launch, B, S, heads, qkv, keymask, ctx, lse, st);
}
int attention_bwd(int dtype, int B, int S, int heads, int dk, const void* qkv, const float* keymask, const void* ctx,
                  const float* lse, const void* d_ctx, void* d_qkv, float* delta, cudaStream_t st, bool allow_split, void* ws,
                  size_t ws_bytes) {
  if (B <= 0 || S <= 0) return MVF_OK;
  MVF_REQUIRE(B <= 65535 && heads <= 65535, MVF_ERR_BAD_ARG, "attention: grid too large");
  if (attention_fa_ok(dtype, S, dk, heads * dk, allow_split, ws, ws_bytes, B, heads) &&
      ((((uintptr_t)qkv) | ((uintptr_t)ctx) | ((uintptr_t)d_ctx) | ((uintptr_t)d_qkv)) & 15) == 0)
    return attention_fa_bwd(B, S, heads, qkv, keymask, ctx, lse, d_ctx, d_qkv, ws, st);
  if (attention_tc_ok(dtype, S, dk, heads * dk, qkv, d_qkv, allow_split) && ((((uintptr_t)ctx) | ((uintptr_t)d_ctx)) & 15) == 0)
    return attention_tc_bwd(B, S, heads, qkv, keymask, ctx, lse, d_ctx, d_qkv, st);
  if (quad_ok(dtype, S, dk, qkv, d_qkv) && (heads * dk) % 4 == 0 && ((((uintptr_t)ctx) | ((uintptr_t)d_ctx)) & 15) == 0) {
    if (dk == 32) return bwd_quad_launch<32>(B, S, heads, qkv, keymask, ctx, lse, d_ctx, d_qkv, st);
    return bwd_quad_launch<64>(B, S, heads, qkv, keymask, ctx, lse, d_ctx, d_qkv, st);
  }
  DISPATCH_DK(bwd_launch, B, S, heads, qkv, keymask, ctx, lse, d_ctx, d_qkv, delta, st);
}

}  // namespace mvf

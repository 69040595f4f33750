// fp32-accumulate SIMT GEMM with arbitrary operand majors.
//
// Role: (1) the fp32 parity path (1e-5 relative needs true fp32 FMA; bf16/tf32 tensor cores cannot give it),
// (2) the on-device cross-check for the tcgen05 GEMM in tests.  It is NOT the throughput path: in bf16 mode
// every contraction goes through gemm_tc.cu.
#include <stdlib.h>

#include "common.cuh"

namespace mvf {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <typename TAB, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(int M, int N, int K, const TAB* __restrict__ A, int64_t sam,
                                                        int64_t sak, const TAB* __restrict__ B, int64_t sbn, int64_t sbk,
                                                        TC* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                                                        const TAB* __restrict__ relu_src, int64_t ld_relu, int flags,
                                                        int a_kmajor, int b_kmajor, int chunk_iters) {
  pdl_entry();
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ty = t / 16, tx = t % 16;
  // Blocked summation: a single fp32 accumulator running over a long K (the weight gradients sum over thousands of
  // entity rows, with heavy cancellation) loses ~K * 2^-24 relative to the partial sums; fp32 partial sums over
  // chunk_iters * BK terms folded into float64 totals keep the parity mode at the 1e-6 level for any K.
  float acc[4][4];
  double tot[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.0; }

  int it = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = t + i * 256;
      int m, k;
      if (a_kmajor) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = to_f<TAB>(A[(int64_t)gm * sam + (int64_t)gk * sak]);
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = t + i * 256;
      int n, k;
      if (b_kmajor) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = to_f<TAB>(B[(int64_t)gn * sbn + (int64_t)gk * sbk]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (chunk_iters > 0 && ++it == chunk_iters) {
      it = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { tot[i][j] += (double)acc[i][j]; acc[i][j] = 0.f; }
    }
    __syncthreads();
  }
  if (chunk_iters > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = (float)(tot[i][j] + (double)acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (flags & MVF_GEMM_RELU) v = fmaxf(v, 0.f);
      if (flags & MVF_GEMM_RELUMASK) v = to_f<TAB>(relu_src[(int64_t)gm * ld_relu + gn]) > 0.f ? v : 0.f;
      int64_t o = (int64_t)gm * ldc + gn;
      if (flags & MVF_GEMM_ACCUM) v += to_f<TC>(C[o]);
      C[o] = from_f<TC>(v);
    }
  }
}

template <typename TAB, typename TC>
static int launch(int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                  int64_t ldb, void* C, int64_t ldc, const float* bias, const void* relu_src, int64_t ld_relu, int flags,
                  cudaStream_t st) {
  dim3 grid(cdiv(N, BN), cdiv(M, BM));
  int64_t sam = a_kmajor ? lda : 1, sak = a_kmajor ? 1 : lda;
  int64_t sbn = b_kmajor ? ldb : 1, sbk = b_kmajor ? 1 : ldb;
  static int chunk = -1;   // MVF_SIMT_CHUNK: BK-steps per fp32 partial sum (0 = one running fp32 sum; A/B measurements)
  if (chunk < 0) {
    const char* e = getenv("MVF_SIMT_CHUNK");
    chunk = e ? atoi(e) : 4;
  }
  launch_k(gemm_simt_kernel<TAB, TC>, grid, 256, 0, st, (int)M, (int)N, (int)K, (const TAB*)A, sam, sak, (const TAB*)B, sbn,
                                                  sbk, (TC*)C, ldc, bias, (const TAB*)relu_src, ld_relu, flags,
                                                  a_kmajor, b_kmajor, K > 4 * BK * (chunk > 0 ? chunk : 1) ? chunk : 0);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

int gemm_simt(int dtype_ab, int dtype_c, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, const void* A,
              int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias, const void* relu_src,
              int64_t ld_relu, int flags, cudaStream_t st) {
  if (M <= 0 || N <= 0) return MVF_OK;
  MVF_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), MVF_ERR_BAD_ARG, "gemm_simt: dims exceed int32");
  if (cdiv(M, BM) > 65535) {
    // grid.y limit: process in row slabs
    int64_t slab = 65535ll * BM;
    for (int64_t m = 0; m < M; m += slab) {
      int64_t mm = (M - m < slab) ? (M - m) : slab;
      size_t esz = dtype_ab == MVF_BF16 ? 2 : 4, csz = dtype_c == MVF_BF16 ? 2 : 4;
      const char* Ap = (const char*)A + (size_t)(a_kmajor ? m * lda : m) * esz;
      char* Cp = (char*)C + (size_t)(m * ldc) * csz;
      const char* Rp = relu_src ? (const char*)relu_src + (size_t)(m * ld_relu) * esz : nullptr;
      MVF_TRY(gemm_simt(dtype_ab, dtype_c, a_kmajor, b_kmajor, mm, N, K, Ap, lda, B, ldb, Cp, ldc, bias, Rp, ld_relu,
                        flags, st));
    }
    return MVF_OK;
  }
  if (dtype_ab == MVF_F32 && dtype_c == MVF_F32)
    return launch<float, float>(a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, bias, relu_src, ld_relu, flags, st);
  if (dtype_ab == MVF_BF16 && dtype_c == MVF_F32)
    return launch<bf16, float>(a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, bias, relu_src, ld_relu, flags, st);
  if (dtype_ab == MVF_BF16 && dtype_c == MVF_BF16)
    return launch<bf16, bf16>(a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, bias, relu_src, ld_relu, flags, st);
  set_error("gemm_simt: unsupported dtype combination ab=%d c=%d", dtype_ab, dtype_c);
  return MVF_ERR_UNSUPPORTED;
}

int gemm_dispatch(int backend, int dtype_ab, int dtype_c, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K,
                  const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
                  const void* relu_src, int64_t ld_relu, int flags, int split_k, cudaStream_t st) {
  // AUTO: bf16 operands -> tcgen05 kind::f16; fp32 operands -> exact fp32 FMA (SIMT).  TCGEN05 with fp32 operands
  // runs them as tf32 on the tensor cores.
  bool use_tc;
  if (backend == MVF_GEMM_SIMT) use_tc = false;
  else if (backend == MVF_GEMM_TCGEN05) use_tc = true;
  else use_tc = (dtype_ab == MVF_BF16);
  if (use_tc) {
    MVF_REQUIRE(dtype_c == MVF_F32 || dtype_ab == MVF_BF16, MVF_ERR_UNSUPPORTED, "tf32 GEMM writes fp32 output only");
    return gemm_tc(dtype_ab, dtype_c, a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, bias, relu_src, ld_relu, flags,
                   split_k, st);
  }
  return gemm_simt(dtype_ab, dtype_c, a_kmajor, b_kmajor, M, N, K, A, lda, B, ldb, C, ldc, bias, relu_src, ld_relu,
                   flags, st);
}

}  // namespace mvf

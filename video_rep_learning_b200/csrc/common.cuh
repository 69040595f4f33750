// Shared device/host helpers for libmvf_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mvf_b200.h"

namespace mvf {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (thread-local message; no exceptions cross the C ABI) ---------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define MVF_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      mvf::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return MVF_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

void count_launch();  // library-wide kernel launch counter (mvf_launch_count)
#define MVF_CHECK_LAUNCH()                  \
  do {                                      \
    mvf::count_launch();                    \
    MVF_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

#define MVF_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      mvf::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

#define MVF_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != MVF_OK) return _s; \
  } while (0)

// ---- programmatic dependent launch ----------------------------------------------------------------------------
// The step is a chain of ~120 short kernels, each depending on its predecessor.  Every kernel is launched with the
// programmatic-stream-serialization attribute and calls pdl_entry() before its first global-memory access:
//   griddepcontrol.wait               blocks until the previous kernel in the stream has completed and flushed its writes;
//   griddepcontrol.launch_dependents  then lets the NEXT kernel's CTAs be scheduled while this one still runs (they park at
//                                     their own wait), so launch latency and kernel prologues leave the critical path.
// Triggering only after the wait keeps the look-ahead at one kernel (no pile-up of parked CTAs).  MVF_PDL=0 launches plainly.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors are picked up by MVF_CHECK_LAUNCH
}
// the same with a thread-block cluster of `cluster_x` CTAs along x (gridDim.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline void launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- scalar conversion ---------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- counter-based dropout ------------------------------------------------------------------------------
// keep(seed, site, idx) is a pure function, so forward and backward regenerate the same mask without
// storing it.  Returns the multiplier (0 or 1/(1-p)).
__host__ __device__ __forceinline__ uint32_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)(x >> 32) ^ (uint32_t)x;
}
__host__ __device__ __forceinline__ float drop_scale(uint64_t seed, int site, uint64_t idx, float p, float inv_keep) {
  uint32_t h = mix64(seed * 0x9E3779B97F4A7C15ULL + ((uint64_t)(site + 1) << 56) + idx);
  float u = (float)(h >> 8) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.0f;
}
// Dropout stream of one step = host value + optional device-resident counter.  The device half exists for CUDA-graph
// replay: kernel arguments are frozen at capture time, so a captured step advances `*dev` on the device (one tiny
// launch per replay) and every replay still draws a fresh mask.  Forward and backward of one step read the same value.
struct DropSeed {
  uint64_t base;
  const uint64_t* dev;
  __host__ __device__ DropSeed(uint64_t b = 0, const uint64_t* d = nullptr) : base(b), dev(d) {}
};
__device__ __forceinline__ float drop_scale(DropSeed s, int site, uint64_t idx, float p, float inv_keep) {
  return drop_scale(s.base + (s.dev ? __ldg(s.dev) : 0ull), site, idx, p, inv_keep);
}

enum DropSite { SITE_FC0 = 0, SITE_POS = 8, SITE_ENC0 = 16 };

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- internal launchers (defined across the .cu files) ---------------------------------------------------
int gemm_simt(int dtype_ab, int dtype_c, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, const void* A,
              int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias, const void* relu_src,
              int64_t ld_relu, int flags, cudaStream_t st);
int gemm_tc(int dtype_ab, int dtype_c, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, const void* A,
            int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias, const void* relu_src,
            int64_t ld_relu, int flags, int split_k, cudaStream_t st);
bool tc_available();

int gemm_dispatch(int backend, int dtype_ab, int dtype_c, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K,
                  const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
                  const void* relu_src, int64_t ld_relu, int flags, int split_k, cudaStream_t st);

}  // namespace mvf

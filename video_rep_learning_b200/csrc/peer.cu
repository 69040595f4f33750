// Cross-rank sums over NVLink / NVSwitch peer memory (one process per GPU, one 8-GPU box).
//
// BatchNorm statistics.  The multi-GPU protocol of the path has one exchange step on the critical path:
// the three train-mode BatchNorms are SyncBatchNorms in the reference (train.py:283), so [sum x, sum x^2] (forward) and
// [sum dy, sum dy*xhat] (backward) -- at most a few KB of float64 -- must be summed over the ranks between two phases of the
// head, six times per step.  An NCCL all-reduce of that size is pure launch + protocol latency.  Here every rank owns a
// symmetric buffer (2 parities x 16 ranks x staging slot, allocated and exchanged by torch's symmetric-memory rendezvous,
// mapped into every peer); one small kernel per rank
//   1. PUSHES its statistics into slot [parity][rank] of every rank's buffer, each float64 as two 8-byte words
//      {32 bits of the value, exchange number}: an 8-byte store arrives whole, so the data carries its own "ready" flag --
//      no fence, no separate flag, one NVLink one-way latency;
//   2. polls the slots of its OWN buffer until every word shows this exchange's number and sums them in rank order --
//      every rank gets the bitwise identical total -- straight into the statistics buffer the next phase reads.
// The exchange counter lives in device memory and is advanced by the kernel, so the launch is CUDA-graph replayable.
// A slot of parity p is rewritten two exchanges later; the writer gets there only after it has received the reader's words of
// the exchange in between, which the reader sends after it has finished reading this one: two parities are enough.
// A rank that never shows up makes the bounded spin trap (sticky CUDA error) instead of hanging the GPU.
#include "kernels.cuh"

namespace mvf {

constexpr int PEER_STAGE_N = 4096;        // float64 values per staging slot (two 8-byte words each)
constexpr int PEER_MAX_WORLD = 16;        // staging slots per parity

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr int PEER_THREADS = 1024;

__global__ void __launch_bounds__(PEER_THREADS)
peer_sum_f64_kernel(double* __restrict__ local, int n, uint8_t* const* __restrict__ bufs, int rank, int world,
                    uint32_t* __restrict__ counter) {
  pdl_entry();
  __shared__ uint32_t s_epoch;
  __shared__ uint8_t* s_buf[64];
  if (threadIdx.x == 0) s_epoch = *counter + 1u;
  if ((int)threadIdx.x < world) s_buf[threadIdx.x] = bufs[threadIdx.x];
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const size_t slot_words = (size_t)2 * PEER_STAGE_N;
  const size_t parity_off = (size_t)(epoch & 1u) * PEER_MAX_WORLD * slot_words;     // in 8-byte words
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t bits = (uint64_t)__double_as_longlong(local[i]);
    const uint64_t w0 = (bits << 32) | epoch, w1 = (bits & 0xffffffff00000000ull) | epoch;
    for (int p = 0; p < world; ++p) {
      uint64_t* dst = reinterpret_cast<uint64_t*>(s_buf[p]) + parity_off + (size_t)rank * slot_words + 2 * i;
      st_relaxed_sys_u64(dst, w0);
      st_relaxed_sys_u64(dst + 1, w1);
    }
  }
  // the words are local now.  All slots of an element are polled together (one L2 round trip per attempt, not one per rank);
  // the additions keep rank order, so every rank computes the bitwise identical total
  const uint64_t* slots = reinterpret_cast<const uint64_t*>(s_buf[rank]) + parity_off;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double t = 0.0;
    for (int p0 = 0; p0 < world; p0 += 8) {
      uint64_t w0[8], w1[8];
      uint32_t spins = 0;
      bool all;
      do {
        all = true;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (p0 + q < world) {
            const uint64_t* src = slots + (size_t)(p0 + q) * slot_words + 2 * i;
            w0[q] = ld_relaxed_sys_u64(src);
            w1[q] = ld_relaxed_sys_u64(src + 1);
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (p0 + q < world) all = all && (uint32_t)w0[q] == epoch && (uint32_t)w1[q] == epoch;
        if (!all && ++spins > (1u << 26)) {
          printf("mvf peer exchange: rank %d timed out in exchange %u\n", rank, epoch);
          __trap();
        }
      } while (!all);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (p0 + q < world) t += __longlong_as_double((long long)((w1[q] & 0xffffffff00000000ull) | (w0[q] >> 32)));
    }
    local[i] = t;
  }
  if (threadIdx.x == 0) *counter = epoch;
}

// ---- gradient all-reduce over the same symmetric memory ---------------------------------------------------------------
// The step ends with one SUM over the ranks of the flat fp32 gradient buffer (19 MB at the Penn shape; DDP's all-reduce,
// train.py:286).  The buffer itself lives in symmetric memory ([gradients | 64 x 64 flags]); every rank reduces ITS
// 1/world slice and broadcasts it:
//   * NVSwitch multicast (mc != null): multimem.ld_reduce pulls the sum of the slice through the switch (the addition happens
//     in the switch, one 16-byte response per element instead of world-1), multimem.st writes it to every rank;
//   * else: world peer loads summed in rank order, world peer stores.
// Each element is added once, by its owner, so every rank ends with the bitwise identical sum.  A handful of CTAs is enough
// (NVLink-latency bound, not SM bound); each CTA does its own two flag barriers with the same-numbered CTA of every peer:
// before the first load (every rank's gradients are final) and after the last store (nobody reads a half-written buffer or
// zeroes a buffer that a peer still reads).  Exchange counters live on the device: CUDA-graph replayable.
constexpr int PEER_AR_CTAS_MAX = 64;
constexpr int PEER_AR_THREADS = 1024;

__device__ __forceinline__ float4 multimem_ld_reduce_f32x4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f32x4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ld_relaxed_sys_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f32x4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flag barrier of CTA blockIdx.x with the same CTA of every peer (flags[b][r] in rank q's buffer: "rank r reached `epoch`")
__device__ __forceinline__ void peer_cta_barrier(uint8_t* const* s_buf, size_t flag_off, int rank, int world, uint32_t epoch) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    st_release_sys(reinterpret_cast<uint32_t*>(s_buf[threadIdx.x] + flag_off) + blockIdx.x * 64 + rank, epoch);
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(s_buf[rank] + flag_off) + blockIdx.x * 64 + threadIdx.x;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
      if (++spins > (1u << 27)) {
        printf("mvf peer all-reduce: rank %d timed out waiting for rank %d (epoch %u)\n", rank, (int)threadIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(PEER_AR_THREADS)
peer_allreduce_f32_kernel(float* __restrict__ mc, uint8_t* const* __restrict__ bufs, size_t data_off, size_t flag_off, int64_t n4,
                          int rank, int world, uint32_t* __restrict__ counters) {
  pdl_entry();
  __shared__ uint8_t* s_buf[64];
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = counters[blockIdx.x];
  if ((int)threadIdx.x < world) s_buf[threadIdx.x] = bufs[threadIdx.x];
  __syncthreads();
  const uint32_t epoch = s_epoch;
  peer_cta_barrier(s_buf, flag_off, rank, world, epoch + 1u);
  const int64_t per = (n4 + world - 1) / world;
  const int64_t lo = (int64_t)rank * per, hi = lo + per < n4 ? lo + per : n4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (mc != nullptr) {
    float* base = mc + data_off / sizeof(float);
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += 8 * stride) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i + u * stride < hi) v[u] = multimem_ld_reduce_f32x4(base + 4 * (i + u * stride));
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i + u * stride < hi) multimem_st_f32x4(base + 4 * (i + u * stride), v[u]);
    }
  } else {
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p0 = 0; p0 < world; p0 += 8) {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (p0 + q < world) v[q] = ld_relaxed_sys_f32x4(reinterpret_cast<const float*>(s_buf[p0 + q] + data_off) + 4 * i);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (p0 + q < world) { t.x += v[q].x; t.y += v[q].y; t.z += v[q].z; t.w += v[q].w; }
      }
      for (int p = 0; p < world; ++p) st_relaxed_sys_f32x4(reinterpret_cast<float*>(s_buf[p] + data_off) + 4 * i, t);
    }
  }
  peer_cta_barrier(s_buf, flag_off, rank, world, epoch + 2u);
  if (threadIdx.x == 0) counters[blockIdx.x] = epoch + 2u;
}

size_t peer_allreduce_flag_bytes() { return (size_t)PEER_AR_CTAS_MAX * 64 * sizeof(uint32_t); }

int peer_allreduce_f32(void* mc_base, void* const* bufs_dev, size_t data_off, size_t flag_off, int64_t n, int rank, int world,
                       uint32_t* counters, int ctas, cudaStream_t st) {
  MVF_REQUIRE(bufs_dev && counters, MVF_ERR_BAD_ARG, "peer_allreduce: null pointer");
  MVF_REQUIRE(n >= 0 && n % 4 == 0 && data_off % 16 == 0 && flag_off % 16 == 0, MVF_ERR_ALIGN,
              "peer_allreduce: %lld floats at offset %zu must be multiples of 16 bytes", (long long)n, data_off);
  MVF_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, MVF_ERR_BAD_ARG, "peer_allreduce: rank %d of %d", rank, world);
  if (n == 0 || world == 1) return MVF_OK;
  if (ctas < 1) ctas = 32;
  if (ctas > PEER_AR_CTAS_MAX) ctas = PEER_AR_CTAS_MAX;
  launch_k(peer_allreduce_f32_kernel, ctas, PEER_AR_THREADS, 0, st, (float*)mc_base, (uint8_t* const*)bufs_dev, data_off, flag_off,
           n / 4, rank, world, counters);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

size_t peer_buffer_bytes() { return (size_t)2 * PEER_MAX_WORLD * PEER_STAGE_N * 16; }

int peer_sum_f64(double* local, int64_t n, void* const* bufs_dev, int rank, int world, uint32_t* counter, cudaStream_t st) {
  MVF_REQUIRE(local && bufs_dev && counter, MVF_ERR_BAD_ARG, "peer_sum: null pointer");
  MVF_REQUIRE(n >= 0 && n <= PEER_STAGE_N, MVF_ERR_UNSUPPORTED, "peer_sum: %lld values > %d", (long long)n, PEER_STAGE_N);
  MVF_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, MVF_ERR_BAD_ARG, "peer_sum: rank %d of %d (at most %d ranks)", rank, world, PEER_MAX_WORLD);
  if (n == 0) return MVF_OK;
  launch_k(peer_sum_f64_kernel, 1, PEER_THREADS, 0, st, local, (int)n, (uint8_t* const*)bufs_dev, rank, world, counter);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace mvf

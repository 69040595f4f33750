// BatchNorm statistics exchange over NVLink / NVSwitch peer memory (one process per GPU, one 8-GPU box).
//
// The multi-GPU protocol of the path has exactly one exchange step on the critical path: the three
// train-mode BatchNorms are SyncBatchNorms in the reference (train.py:283), so [sum x, sum x^2] (forward) and
// [sum dy, sum dy*xhat] (backward) -- at most a few KB of float64 -- must be summed over the ranks between two phases
// of the head, eight times per step.  An NCCL all-reduce of that size is pure launch + protocol latency.  Here every rank
// owns a symmetric buffer ([flags | 2 staging slots], allocated and exchanged by torch's symmetric-memory rendezvous,
// mapped into every peer); one small kernel per rank
//   1. copies its statistics into its own staging slot (parity = exchange counter & 1),
//   2. publishes a flag in every peer's buffer (st.release.sys over NVLink) and waits for the flags of all peers,
//   3. reads all staging slots over NVLink and sums them in rank order -- every rank gets the bitwise identical total --
//      straight into the statistics buffer the next phase reads.
// The exchange counter lives in device memory and is advanced by the kernel, so the launch is CUDA-graph replayable.
// A slot of parity p is rewritten two exchanges later, after every rank has passed the flag wait of the exchange in
// between, which it only reaches after finishing its reads: no second barrier is needed.  A rank that never shows up
// makes the bounded spin trap (sticky CUDA error) instead of hanging the GPU.
#include "kernels.cuh"

namespace mvf {

constexpr int PEER_FLAG_BYTES = 256;      // 64 x uint32 flags at the start of every rank's buffer
constexpr int PEER_STAGE_N = 4096;        // doubles per staging slot

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
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
  const size_t slot = PEER_FLAG_BYTES + (size_t)(epoch & 1u) * PEER_STAGE_N * sizeof(double);
  double* mine = reinterpret_cast<double*>(s_buf[rank] + slot);
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = local[i];
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    // flag [rank] in peer threadIdx.x's buffer: "rank's slot of this exchange is complete"
    st_release_sys(reinterpret_cast<uint32_t*>(s_buf[threadIdx.x]) + rank, epoch);
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(s_buf[rank]) + threadIdx.x;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
      if (++spins > (1u << 27)) {
        printf("mvf peer exchange: rank %d timed out waiting for rank %d (exchange %u)\n", rank, (int)threadIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  // all peer loads of an element are issued before the first addition (eight NVLink round trips in flight instead of
  // one after the other); the additions keep rank order, so every rank computes the bitwise identical total
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double t = 0.0;
    for (int p0 = 0; p0 < world; p0 += 8) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        v[q] = (p0 + q < world) ? ld_relaxed_sys_f64(reinterpret_cast<const double*>(s_buf[p0 + q] + slot) + i) : 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (p0 + q < world) t += v[q];
    }
    local[i] = t;
  }
  if (threadIdx.x == 0) *counter = epoch;
}

size_t peer_buffer_bytes() { return PEER_FLAG_BYTES + (size_t)2 * PEER_STAGE_N * sizeof(double); }

int peer_sum_f64(double* local, int64_t n, void* const* bufs_dev, int rank, int world, uint32_t* counter, cudaStream_t st) {
  MVF_REQUIRE(local && bufs_dev && counter, MVF_ERR_BAD_ARG, "peer_sum: null pointer");
  MVF_REQUIRE(n >= 0 && n <= PEER_STAGE_N, MVF_ERR_UNSUPPORTED, "peer_sum: %lld values > %d", (long long)n, PEER_STAGE_N);
  MVF_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, MVF_ERR_BAD_ARG, "peer_sum: rank %d of %d", rank, world);
  if (n == 0) return MVF_OK;
  launch_k(peer_sum_f64_kernel, 1, PEER_THREADS, 0, st, local, (int)n, (uint8_t* const*)bufs_dev, rank, world, counter);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace mvf

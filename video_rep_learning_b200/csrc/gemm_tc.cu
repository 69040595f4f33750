// bf16 x bf16 -> fp32 (kind::f16) and tf32 x tf32 -> fp32 (kind::tf32, fp32 operands in memory) GEMM on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA (cp.async.bulk.tensor,
// 128B swizzle) through an mbarrier ring.  sm_100a only.  The two element types share one kernel: a stage row is
// always 128 bytes of K (64 bf16 or 32 fp32) and one MMA consumes 32 bytes of K (UMMA_K = 16 or 8).
//
//   C[M,N] (+)= opA(A)[M,K] * opB(B)[K,N] (+ bias)           fp32 accumulate in TMEM
//
// Operand majors are free: K-major ([rows,K] row-major, the nn.Linear forward) or MN-major ([K,rows]
// row-major, needed by dX = dY*W and dW = dY^T*X) -- selected per operand through the shared-memory
// descriptor + instruction descriptor, never by materialising a transpose.
//
// Kernel shape: persistent, one CTA per SM, 12 warps (16 in SPLIT3 mode):
//   warp 0      TMA producer (one elected lane)
//   warp 1      MMA issuer   (one elected lane; tcgen05.mma cta_group::1, M=128, N=BLOCK_N, K=16)
//   warp 2      TMEM allocator / deallocator
//   warps 4..7 + the last four warps   epilogue, two warps per TMEM lane quadrant (= warp%4), each taking half of the
//               tile's columns: tcgen05.ld -> bias/ReLU/mask -> swizzled staging tile -> TMA store / reduce-add (fp32)
//               or registers -> global (bf16)
// Two TMEM accumulators (2*BLOCK_N columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// Split-K (work item = tile x K-slice, fp32 atomics) fills the machine when M*N is small and K huge
// (the weight gradient of the K/V projection: 768 x 2304 outputs, K = frames*196).
//
// SPLIT3 ("bf16x3") mode for fp32 operands, both K-major: kind::tf32 truncates the operands to 10 mantissa bits,
// which the 1/temperature of SCL amplifies to a few 1e-2 on the gradients.  In this mode four extra warps (8..11)
// rewrite every landed fp32 stage IN PLACE as bf16 pairs -- a 128-byte row of 32 fp32 becomes [hi(32) | lo(32)] bf16
// with hi = bf16(x), lo = bf16(x - hi), same swizzle -- and the MMA warp issues A_lo*B_hi + A_hi*B_lo + A_hi*B_hi on
// kind::f16 (6 MMAs of K=16 per stage).  16 mantissa bits per operand at 1.5x the tensor time of tf32.
#include <stdlib.h>

#include "tc_common.cuh"

namespace mvf {

namespace tc {

struct Params {
  int M, N, K;
  int tiles_m, tiles_n, split_k, kb_per_split, num_kb;
  int a_kmajor, b_kmajor;
  int c_bf16;
  int b_presplit;  // SPLIT3: the B tile lands already converted (MVF_GEMM_B_PRESPLIT): only the A rows are rewritten
  int c_tma;   // fp32 C written by TMA (store, or reduce-add for ACCUM / split-K) from a swizzled shared-memory staging tile
  int flags;
  void* C;
  int64_t ldc;
  const float* bias;
  const void* relu_src;  // same element type as the operands
  int64_t ld_relu;
  unsigned long long* dbg;  // MVF_GEMM_DBG=1: %globaltimer stamps of CTA 0 (debug aid, null otherwise)
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define MVF_STAMP(i) do { if (p.dbg != nullptr && blockIdx.x == 0) p.dbg[i] = gtimer(); } while (0)

template <int BLOCK_N, int STAGES, int EB, bool SPLIT>
__global__ void __launch_bounds__((SPLIT ? NUM_THREADS_SPLIT : NUM_THREADS) + EXTRA_EPI_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const Params p) {
  constexpr int BLOCK_K = ROW_BYTES / EB;   // elements of K per stage (64 bf16 / 32 tf32)
  constexpr int UMMA_K = 32 / EB;           // elements of K per tcgen05.mma (16 / 8)
  constexpr int CHUNK = ROW_BYTES / EB;     // M/N elements per 128-byte row of an MN-major operand
  constexpr int A_BYTES = BLOCK_M * ROW_BYTES;
  constexpr int B_BYTES = BLOCK_N * ROW_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                            : (2 * BLOCK_N <= 256) ? 256 : 512;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A|B)] then barriers
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* conv_bar = tmem_empty + 2;  // SPLIT: stage rewritten as bf16 hi|lo, ready for the MMA warp
  uint32_t* tmem_slot = (uint32_t*)(conv_bar + STAGES);
  // epilogue staging for the TMA store path: 4 warps x 2 buffers x [32 rows x 128 B], 1024-byte aligned
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES + 1024;
  static_assert(!SPLIT || EB == 4, "SPLIT3 applies to fp32 operands");
  static_assert((3 * STAGES + 4) * 8 + 4 <= 256, "barrier block overflows its 256 bytes");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_work = p.tiles_m * p.tiles_n * p.split_k;
  if (threadIdx.x == 0) MVF_STAMP(0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    if (p.c_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&conv_bar[s], NUM_THREADS_SPLIT - NUM_THREADS);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) MVF_STAMP(1);
  // everything above touched only shared / tensor memory and the kernel parameters: with programmatic dependent launch it
  // overlaps the tail of the previous kernel in the stream; global memory is first read / written below
  pdl_entry();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int split = w / (p.tiles_m * p.tiles_n);
        const int rem = w - split * (p.tiles_m * p.tiles_n);
        const int tm = rem / p.tiles_n, tn = rem - tm * p.tiles_n;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k = kb * BLOCK_K;
          if (p.a_kmajor) {
            tma_load_2d(&map_a, &full_bar[stage], sa, k, tm * BLOCK_M);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_M / CHUNK; ++c)
              tma_load_2d(&map_a, &full_bar[stage], sa + c * (BLOCK_K * 128), tm * BLOCK_M + c * CHUNK, k);
          }
          if (p.b_kmajor) {
            tma_load_2d(&map_b, &full_bar[stage], sb, k, tn * BLOCK_N);
          } else {
#pragma unroll
            for (int c = 0; c < BLOCK_N / CHUNK; ++c)
              tma_load_2d(&map_b, &full_bar[stage], sb + c * (BLOCK_K * 128), tn * BLOCK_N + c * CHUNK, k);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Every lane runs the protocol with warp-uniform operands and the elected lane issues (umma_p): under a divergent
    // `if (lane == 0)` every tcgen05.mma is wrapped in an ELECT / R2UR.BROADCAST / branch loop of ~100 clocks, as long as
    // the tensor time of an N = 128 MMA -- it dominates the short GEMMs of the chain.
    {
      const uint32_t leader = elect_leader();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0), smem_a0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t idesc = SPLIT ? make_idesc(BLOCK_N, true, true, 2) : make_idesc(BLOCK_N, p.a_kmajor != 0, p.b_kmajor != 0, EB);
      // start-address advance (in 16 B units) per UMMA_K step inside a stage: 32 B along a K-major row,
      // UMMA_K rows of 128 B for an MN-major operand
      const uint32_t adv_a = p.a_kmajor ? (UMMA_K * EB) >> 4 : (UMMA_K * 128) >> 4;
      const uint32_t adv_b = p.b_kmajor ? (UMMA_K * EB) >> 4 : (UMMA_K * 128) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int split = w / (p.tiles_m * p.tiles_n);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 2);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_u + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(SPLIT ? &conv_bar[stage] : &full_bar[stage], phase, 3);
          tcgen05_fence_after();
          if (it == 0 && kb == kb0 && lane == 0) MVF_STAMP(2);
          const uint32_t sa = smem_a0 + stage * STAGE_BYTES;
          if constexpr (SPLIT) {
            // row = [hi k0..15 | hi k16..31 | lo k0..15 | lo k16..31] bf16, 32 bytes each: +2 descriptor units per slot
            const uint64_t da = make_smem_desc(sa, true, BLOCK_K, 2);
            const uint64_t db = make_smem_desc(sa + A_BYTES, true, BLOCK_K, 2);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t hi = (uint64_t)(2 * j), lo = (uint64_t)(4 + 2 * j);
              umma_p<2>(tmem_d, da + lo, db + hi, idesc, (kb > kb0 || j > 0) ? 1u : 0u, leader);   // small terms first
              umma_p<2>(tmem_d, da + hi, db + lo, idesc, 1u, leader);
              umma_p<2>(tmem_d, da + hi, db + hi, idesc, 1u, leader);
            }
          } else {
            const uint64_t da = make_smem_desc(sa, p.a_kmajor != 0, BLOCK_K, EB);
            const uint64_t db = make_smem_desc(sa + A_BYTES, p.b_kmajor != 0, BLOCK_K, EB);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_p<EB>(tmem_d, da + (uint64_t)(k * adv_a), db + (uint64_t)(k * adv_b), idesc, (kb > kb0 || k > 0) ? 1u : 0u, leader);
            }
          }
          umma_commit_p(&empty_bar[stage], leader);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_p(&tmem_full[acc], leader);  // accumulator complete -> epilogue
        if (it == 0 && lane == 0) MVF_STAMP(3);
      }
    }
  } else if (SPLIT && warp >= 8 && warp < 12) {
    // ===================== fp32 -> bf16 hi|lo converter (SPLIT3) =====================
    // Thread t owns rows t, t+128, ... of the stage (A rows then B rows; both tiles are K-major, 128-byte rows,
    // SWIZZLE_128B: 16-byte chunk c of row r lives at chunk c ^ (r & 7)).  It reads its whole row and writes it back,
    // so the rewrite is race-free in place; fence.proxy.async makes it visible to the tensor core's async proxy.
    const int ct = threadIdx.x - NUM_THREADS;
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int split = w / (p.tiles_m * p.tiles_n);
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase, 5);
        uint8_t* tile = smem + stage * STAGE_BYTES;
        const int conv_rows = p.b_presplit ? BLOCK_M : BLOCK_M + BLOCK_N;
#pragma unroll 1
        for (int r = ct; r < conv_rows; r += NUM_THREADS_SPLIT - NUM_THREADS) {
          const uint32_t rowa = smem_u32(tile + r * ROW_BYTES);
          const int sw = r & 7;
          uint4 v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = lds128(rowa + ((c ^ sw) << 4));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float x[8] = {__uint_as_float(v[2 * c].x), __uint_as_float(v[2 * c].y), __uint_as_float(v[2 * c].z),
                                __uint_as_float(v[2 * c].w), __uint_as_float(v[2 * c + 1].x), __uint_as_float(v[2 * c + 1].y),
                                __uint_as_float(v[2 * c + 1].z), __uint_as_float(v[2 * c + 1].w)};
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
              const float2 hf = __bfloat1622float2(h);
              const __nv_bfloat162 l = __floats2bfloat162_rn(x[2 * j] - hf.x, x[2 * j + 1] - hf.y);
              hi[j] = *reinterpret_cast<const uint32_t*>(&h);
              lo[j] = *reinterpret_cast<const uint32_t*>(&l);
            }
            sts128(rowa + ((c ^ sw) << 4), make_uint4(hi[0], hi[1], hi[2], hi[3]));
            sts128(rowa + (((4 + c) ^ sw) << 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&conv_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if ((warp >= 4 && warp < 8) || warp >= (SPLIT ? 12 : 8)) {
    // ===================== epilogue =====================
    // two warps per TMEM lane quadrant: `half` 0 (warps 4..7) takes the first BLOCK_N/2 columns of the tile, `half` 1
    // (the last four warps of the CTA) the rest
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int half = warp >= 8 ? 1 : 0;
    constexpr int C_BEGIN1 = (BLOCK_N / 2 + 31) / 32 * 32;   // first column chunk of half 1
    const int c_begin = half ? C_BEGIN1 : 0, c_end = half ? BLOCK_N : C_BEGIN1;
    int it = 0;
    uint32_t n_chunks = 0;   // TMA path: chunks issued by this warp
    uint8_t* my_stage = epi_stage + (half * 4 + q) * 4096;   // one 32 x 128 B staging tile per epilogue warp
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int split = w / (p.tiles_m * p.tiles_n);
      const int rem = w - split * (p.tiles_m * p.tiles_n);
      const int tm = rem / p.tiles_n, tn = rem - tm * p.tiles_n;
      const int kb0 = split * p.kb_per_split;
      const bool has_k = kb0 < p.num_kb;
      const int row = tm * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M;
      const bool add_bias = p.bias != nullptr && split == 0;
      const bool atomic = p.split_k > 1;
      if (p.c_tma) {
        // ---- fp32 output through TMA: registers -> swizzled staging tile -> cp.async.bulk.tensor store / reduce-add ----
        const int row0 = tm * BLOCK_M + q * 32;
        const bool reduce = atomic || (p.flags & MVF_GEMM_ACCUM);
        // the bias of the first chunk is fetched before the accumulator is waited for (overlaps the main loop)
        float bnext = 0.f;
        {
          const int c = tn * BLOCK_N + c_begin + lane;
          if (add_bias && c < p.N) bnext = __ldg(p.bias + c);
        }
        mbar_wait(&tmem_full[acc], acc_phase, 4);
        tcgen05_fence_after();
        if (it == 0 && warp == 4 && lane == 0) MVF_STAMP(4);
        if (has_k && row0 < p.M) {
#pragma unroll 1
          for (int c0 = c_begin; c0 < c_end; c0 += 32) {
            const int col0 = tn * BLOCK_N + c0;
            if (col0 >= p.N) break;  // warp-uniform
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), r);
            const float bcur = bnext;
            {
              const int c = col0 + 32 + lane;
              bnext = (add_bias && c0 + 32 < c_end && c < p.N) ? __ldg(p.bias + c) : 0.f;
            }
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (add_bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bcur, j);
            }
            if (p.flags & MVF_GEMM_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if ((p.flags & MVF_GEMM_RELUMASK) && row_ok) {
              if constexpr (EB == 2) {
                const bf16* rs = (const bf16*)p.relu_src + (int64_t)row * p.ld_relu + col0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) v[j] = (__bfloat162float(rs[j]) > 0.f) ? v[j] : 0.f;
              } else {
                const float* rs = (const float*)p.relu_src + (int64_t)row * p.ld_relu + col0;
                if (col0 + 32 <= p.N && (p.ld_relu & 3) == 0 && ((((uintptr_t)p.relu_src) & 15) == 0)) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    const float4 m = *reinterpret_cast<const float4*>(rs + j);
                    v[j] = m.x > 0.f ? v[j] : 0.f; v[j + 1] = m.y > 0.f ? v[j + 1] : 0.f;
                    v[j + 2] = m.z > 0.f ? v[j + 2] : 0.f; v[j + 3] = m.w > 0.f ? v[j + 3] : 0.f;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (col0 + j < p.N) v[j] = (rs[j] > 0.f) ? v[j] : 0.f;
                }
              }
            }
            uint8_t* sb = my_stage;
            if (n_chunks >= 1) {   // the previous store of this warp must have drained the staging tile
              if (lane == 0) bulk_wait_read_0();
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(smem_u32(sb) + lane * 128 + ((j ^ (lane & 7)) << 4),
                     make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                __float_as_uint(v[4 * j + 3])));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (reduce) tma_reduce_add_2d(&map_c, sb, col0, row0);
              else tma_store_2d(&map_c, sb, col0, row0);
              bulk_commit();
            }
            ++n_chunks;
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (it == 0 && warp == 4 && lane == 0) MVF_STAMP(5);
        continue;
      }
      mbar_wait(&tmem_full[acc], acc_phase, 4);
      tcgen05_fence_after();
      if (it == 0 && warp == 4 && lane == 0) MVF_STAMP(4);
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        const int col0 = tn * BLOCK_N + c0;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), r);
        tmem_ld_wait();
        if (row_ok && has_k) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (add_bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
          }
          if (p.flags & MVF_GEMM_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.flags & MVF_GEMM_RELUMASK) {
            if constexpr (EB == 2) {
              const bf16* rs = (const bf16*)p.relu_src + (int64_t)row * p.ld_relu + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] = (__bfloat162float(rs[j]) > 0.f) ? v[j] : 0.f;
            } else {
              const float* rs = (const float*)p.relu_src + (int64_t)row * p.ld_relu + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] = (rs[j] > 0.f) ? v[j] : 0.f;
            }
          }
          const bool full = col0 + 32 <= p.N;
          if (p.c_bf16) {
            bf16* cp = (bf16*)p.C + (int64_t)row * p.ldc + col0;
            if (full && ((p.ldc & 7) == 0) && ((((uintptr_t)p.C) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2);
                pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(cp + j) = pk;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) cp[j] = __float2bfloat16_rn(v[j]);
            }
          } else {
            // fp32 output: each thread owns 128 contiguous bytes of its row (8 x 16-byte stores; measured 25x fewer
            // instructions than a shared-memory transpose to row-coalesced 4-byte stores, profiles/r01_gemm_small.txt)
            float* cp = (float*)p.C + (int64_t)row * p.ldc + col0;
            if (atomic) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) atomicAdd(cp + j, v[j]);
            } else if (p.flags & MVF_GEMM_ACCUM) {
              if (full && ((p.ldc & 3) == 0) && ((((uintptr_t)p.C) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 o = *reinterpret_cast<float4*>(cp + j);
                  o.x += v[j]; o.y += v[j + 1]; o.z += v[j + 2]; o.w += v[j + 3];
                  *reinterpret_cast<float4*>(cp + j) = o;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) cp[j] += v[j];
              }
            } else if (full && ((p.ldc & 3) == 0) && ((((uintptr_t)p.C) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) cp[j] = v[j];
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (it == 0 && warp == 4 && lane == 0) MVF_STAMP(5);
    }
    if (p.c_tma && lane == 0) bulk_wait_all();   // every TMA store of this warp has completed before the CTA exits
  }

  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) MVF_STAMP(6);
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------
// rows x cols row-major matrix (bf16 when eb == 2, fp32 when eb == 4) with leading dimension ld (elements);
// box = box_cols x box_rows.
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                    int box_rows, int eb, bool mn_major) {
  EncodeTiledFn fn = get_encode_fn();
  MVF_REQUIRE(fn != nullptr, MVF_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  MVF_REQUIRE((((uintptr_t)ptr) & 15) == 0, MVF_ERR_ALIGN, "gemm_tc: operand base %p not 16-byte aligned", ptr);
  MVF_REQUIRE((ld * eb) % 16 == 0, MVF_ERR_ALIGN, "gemm_tc: leading dimension %lld not a multiple of %d elements",
              (long long)ld, 16 / eb);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * eb};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  (mn_major && eb == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MVF_REQUIRE(r == CUDA_SUCCESS, MVF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
              (long long)rows, (long long)cols, (long long)ld);
  return MVF_OK;
}

template <int BLOCK_N, int STAGES, int EB, bool SPLIT = false>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const Params& p, int num_sms,
                  cudaStream_t st) {
  // stages | 1 KB (barriers + TMEM slot) | 32 KB epilogue staging | 1 KB alignment slack
  constexpr int smem = STAGES * (BLOCK_M * ROW_BYTES + BLOCK_N * ROW_BYTES) + 1024 + 4 * 8192 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    MVF_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, STAGES, EB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        smem));
    configured = true;
  }
  int work = p.tiles_m * p.tiles_n * p.split_k;
  int grid = work < num_sms ? work : num_sms;
  launch_k(gemm_tc_kernel<BLOCK_N, STAGES, EB, SPLIT>, grid, (SPLIT ? NUM_THREADS_SPLIT : NUM_THREADS) + EXTRA_EPI_THREADS, smem, st, ma, mb, mc, p);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace tc

static int g_num_sms = -1;
static int g_cc_major = -1;
static void query_device() {
  if (g_num_sms >= 0) return;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
    g_num_sms = prop.multiProcessorCount;
    g_cc_major = prop.major;
  } else {
    g_num_sms = 0;
    g_cc_major = 0;
  }
}

bool tc_available() {
  query_device();
  return g_cc_major == 10 && tc::get_encode_fn() != nullptr;
}

int gemm_tc(int dtype_ab, int dtype_c, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, const void* A,
            int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias, const void* relu_src,
            int64_t ld_relu, int flags, int split_k, cudaStream_t st) {
  using namespace tc;
  const int eb = dtype_ab == MVF_BF16 ? 2 : 4;   // fp32 operands run as tf32 (10-bit mantissa) on the tensor cores
  const int BLOCK_K = ROW_BYTES / eb;
  const int CHUNK = ROW_BYTES / eb;
  if (M <= 0 || N <= 0) return MVF_OK;
  MVF_REQUIRE(tc_available(), MVF_ERR_UNSUPPORTED, "tcgen05 GEMM requested but the device is not sm_100");
  MVF_REQUIRE(K > 0, MVF_ERR_BAD_ARG, "gemm_tc: K must be positive");
  MVF_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), MVF_ERR_BAD_ARG, "gemm_tc: dims exceed int32");
  int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  // small problems: trade tile size for CTA count so that the 148 SMs are not left idle
  {
    const int kb_est = cdiv(K, BLOCK_K);
    int sk_est = split_k > 0 ? split_k : ((dtype_c == MVF_F32 && !(flags & (MVF_GEMM_RELU | MVF_GEMM_RELUMASK)) && kb_est >= 16) ? kb_est / 8 : 1);
    if (sk_est < 1) sk_est = 1;
    static int fill_target = -1;  // tuning knob: CTAs wanted before tiles stop shrinking
    if (fill_target < 0) {
      const char* e = getenv("MVF_GEMM_FILL");
      fill_target = e ? atoi(e) : 60;  // measured best with the side stream (profiles/r01_ab_knobs.txt)
    }
    // Measured in round 2 and not adopted (cfg2 step, 1.508 ms): smaller tiles for the forward chain alone -- it has no
    // side-stream GEMM beside it -- 120 CTAs no change, 148 / 240 CTAs +1.3 % / +1.6 %; eight converter warps (two threads
    // per stage row) +0 % in the step, +10 % on the stand-alone SPLIT3 GEMM; the A operand pre-split by a separate
    // full-machine kernel (converter warps idle): +2.8 % -- the extra launch costs more than the ~2.5 us it takes off a GEMM;
    // no split-K (and no output memset) for the K = 512 GEMMs: -0.3 %, within run-to-run noise.
    while (bn > 64 && (int64_t)cdiv(M, BLOCK_M) * cdiv(N, bn) * sk_est < fill_target) bn >>= 1;
  }
  Params p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.tiles_m = cdiv(M, BLOCK_M);
  p.tiles_n = cdiv(N, bn);
  p.num_kb = cdiv(K, BLOCK_K);
  if (split_k <= 0) {
    // auto: split K only when the output tiles cannot fill the machine and K is long
    int tiles = p.tiles_m * p.tiles_n;
    split_k = 1;
    if (dtype_c == MVF_F32 && !(flags & (MVF_GEMM_RELU | MVF_GEMM_RELUMASK)) && tiles < g_num_sms && p.num_kb >= 16) {
      split_k = (2 * g_num_sms + tiles - 1) / tiles;
      int max_split = p.num_kb / 8;
      if (split_k > max_split) split_k = max_split;
      if (split_k < 1) split_k = 1;
    }
  }
  if (split_k > 1) {
    MVF_REQUIRE(dtype_c == MVF_F32 && !(flags & (MVF_GEMM_RELU | MVF_GEMM_RELUMASK)), MVF_ERR_BAD_ARG,
                "gemm_tc: split-K needs an fp32 linear epilogue");
    if (!(flags & MVF_GEMM_ACCUM)) {
      // partial sums are accumulated with atomics: C must start from zero
      MVF_CHECK_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
    }
  }
  p.kb_per_split = cdiv(p.num_kb, split_k);
  p.split_k = cdiv(p.num_kb, p.kb_per_split);
  p.a_kmajor = a_kmajor; p.b_kmajor = b_kmajor;
  p.c_bf16 = dtype_c == MVF_BF16;
  const bool split3 = eb == 4 && (flags & MVF_GEMM_SPLIT3) && a_kmajor && b_kmajor;
  p.b_presplit = (split3 && (flags & MVF_GEMM_B_PRESPLIT)) ? 1 : 0;
  MVF_REQUIRE(!(flags & MVF_GEMM_B_PRESPLIT) || split3, MVF_ERR_BAD_ARG,
              "gemm_tc: a pre-split B operand needs SPLIT3 with fp32, K-major operands");
  MVF_REQUIRE(!p.b_presplit || ldb % 32 == 0, MVF_ERR_ALIGN, "gemm_tc: pre-split B needs ldb (%lld) to be a multiple of 32",
              (long long)ldb);
  p.flags = flags;
  p.C = C; p.ldc = ldc; p.bias = bias;
  p.relu_src = relu_src; p.ld_relu = ld_relu;
  p.dbg = nullptr;
  static int dbg_on = -1;
  static unsigned long long* dbg_buf = nullptr;
  if (dbg_on < 0) {
    const char* e = getenv("MVF_GEMM_DBG");
    dbg_on = (e && atoi(e) != 0) ? 1 : 0;
    if (dbg_on && cudaMalloc(&dbg_buf, 8 * sizeof(unsigned long long)) != cudaSuccess) dbg_on = 0;   // debug aid only
  }
  if (dbg_on) {
    cudaMemsetAsync(dbg_buf, 0, 8 * sizeof(unsigned long long), st);
    p.dbg = dbg_buf;
  }

  CUtensorMap ma, mb, mc;
  memset(&mc, 0, sizeof(mc));
  // fp32 C goes out through TMA whenever its rows are 16-byte aligned (MVF_GEMM_TMA_STORE=0 keeps the register path)
  static int tma_store_on = -1;
  if (tma_store_on < 0) {
    const char* e = getenv("MVF_GEMM_TMA_STORE");
    tma_store_on = (e && atoi(e) == 0) ? 0 : 1;
  }
  p.c_tma = (tma_store_on && dtype_c == MVF_F32 && (ldc % 4) == 0 && ((((uintptr_t)C) & 15) == 0)) ? 1 : 0;
  if (p.c_tma) MVF_TRY(make_map(&mc, C, M, N, ldc, 32, 32, 4, false));
  if (a_kmajor) MVF_TRY(make_map(&ma, A, M, K, lda, BLOCK_K, BLOCK_M, eb, false));
  else MVF_TRY(make_map(&ma, A, K, M, lda, CHUNK, BLOCK_K, eb, true));
  if (b_kmajor) MVF_TRY(make_map(&mb, B, N, p.b_presplit ? round_up(K, 32) : K, ldb, BLOCK_K, bn, eb, false));
  else MVF_TRY(make_map(&mb, B, K, N, ldb, CHUNK, BLOCK_K, eb, true));

  int rc;
  if (eb == 2) {
    if (bn == 256) rc = launch<256, 4, 2>(ma, mb, mc, p, g_num_sms, st);
    else if (bn == 128) rc = launch<128, 6, 2>(ma, mb, mc, p, g_num_sms, st);
    else rc = launch<64, 8, 2>(ma, mb, mc, p, g_num_sms, st);
  } else if ((flags & MVF_GEMM_SPLIT3) && a_kmajor && b_kmajor) {
    if (bn == 256) rc = launch<256, 4, 4, true>(ma, mb, mc, p, g_num_sms, st);
    else if (bn == 128) rc = launch<128, 6, 4, true>(ma, mb, mc, p, g_num_sms, st);
    else rc = launch<64, 8, 4, true>(ma, mb, mc, p, g_num_sms, st);
  } else {
    if (bn == 256) rc = launch<256, 4, 4>(ma, mb, mc, p, g_num_sms, st);
    else if (bn == 128) rc = launch<128, 6, 4>(ma, mb, mc, p, g_num_sms, st);
    else rc = launch<64, 8, 4>(ma, mb, mc, p, g_num_sms, st);
  }
  if (rc == MVF_OK && dbg_on) {
    unsigned long long h[8];
    if (cudaStreamSynchronize(st) == cudaSuccess && cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
      auto d = [&](int i) { return h[i] >= h[0] ? (double)(h[i] - h[0]) * 1e-3 : -1.0; };
      fprintf(stderr, "gemm_tc dbg M=%lld N=%lld K=%lld bn=%d split=%d eb=%d: setup %.2f | first stage ready %.2f | mma issued %.2f | "
              "acc ready %.2f | epilogue done %.2f | exit %.2f us\n", (long long)M, (long long)N, (long long)K, bn, p.split_k, eb,
              d(1), d(2), d(3), d(4), d(5), d(6));
    }
  }
  return rc;
}

}  // namespace mvf

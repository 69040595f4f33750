// Sequence Contrastive Loss, forward + gradient fused (algos/scl.py:52-105).
//
// The reference builds ~10 N x N fp32 temporaries (N = Bv*2*T) with Python loops over the batch.  Here:
//   * the useful work is per video pair: one T x T logit block S = E0 E1^T / tau serves both view directions;
//   * a row i = (video, view, frame) only interacts with its partner block (own video, other view), plus
//     the "extra" columns that enter its partition sum Z_i:
//       - every MASKED frame of the local batch with weight 1e-6 (weight.masked_fill_(mask==0, 1e-6) runs
//         after the single/noself zeroing, scl.py:80)                                   [quirk = 1]
//       - for NEGATIVE_TYPE batch_noself every valid frame of the OTHER videos with weight 1 (scl.py:74-79).
// Launch sequence (all on one stream, no host sync; M = sum(mask) stays on the device):
//   prep -> cross(sum only: Z extras) -> rowstats (Z, g, loss) -> grad (own-pair dE) -> cross(accumulate dE extras)
// Embeddings are staged in shared memory tiles, rows are owned by warps (lanes = partner columns), row
// reductions are warp shuffles; nothing of size T x T or N x N ever reaches HBM.
//
// Closed form, direction with rows i and partner columns j, mm_ij = m_i m_j:
//   y_ij = pw_ij / sum_j pw_ij,  pw_ij = mm_ij exp(-d_ij^2 / (2 var)),  d_ij = |fl(fl(s_i / L_i) * L_j) - s_j|
//   Z_i = sum_j mm_ij e^{l_ij} + zext_i,  p = e^{l}/Z,  q = p + 1e-6,  loss += mm y (log y - log q) / M
//   r = p/q,  g_i = sum_j mm y r,  dloss/dl_ij = mm (p g_i - y r) / M,  extras: w c_i e^{l_ik},  c_i = g_i/(Z_i M)
#include "kernels.cuh"

namespace mvf {

constexpr int SCL_MAXD = 256;  // embedding width limit of the register d-slices (8 per lane)
constexpr int SCL_MAXTC = 8;   // ceil(T/32) limit -> T <= 256

struct SclWs {
  float* M;       // [1] sum of masks
  float* Z;       // [N]
  float* g;       // [N]
  float* den;     // [N]
  float* c;       // [N]  g / (Z M)
  float* zext;    // [N]
  int* counts;    // [2] n_valid, n_masked
  int* valid;     // [N]
  int* masked;    // [N]
  int* chunk;     // [2 * (nchunks + 1)] valid / masked counts per 1024-row chunk, then their exclusive scans (large N only)
  float* mex;     // [n_valid x n_masked] 1e-6 e^{l_rk} of (valid row, masked column) pairs (N <= SCL_MEX_MAXN only, else null)
};
constexpr int SCL_MEX_MAXN = 4096;   // the stored matrix has at most N^2 / 4 entries: 16.8 MB at this N

static size_t scl_ws_layout(int N, SclWs* w, char* base) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += (bytes + 255) / 256 * 256;
    return p;
  };
  float* M = (float*)take(sizeof(float) * 4);
  float* Z = (float*)take(sizeof(float) * N);
  float* g = (float*)take(sizeof(float) * N);
  float* den = (float*)take(sizeof(float) * N);
  float* c = (float*)take(sizeof(float) * N);
  float* zext = (float*)take(sizeof(float) * N);
  int* counts = (int*)take(sizeof(int) * 4);
  int* valid = (int*)take(sizeof(int) * N);
  int* masked = (int*)take(sizeof(int) * N);
  int* chunk = (int*)take(sizeof(int) * 2 * ((size_t)(N + 1023) / 1024 + 1));
  float* mex = N <= SCL_MEX_MAXN ? (float*)take(sizeof(float) * ((size_t)N * N / 4 + 1)) : nullptr;
  if (w) { w->chunk = chunk; w->mex = mex; }
  if (w) { w->M = M; w->Z = Z; w->g = g; w->den = den; w->c = c; w->zext = zext; w->counts = counts; w->valid = valid; w->masked = masked; }
  return off;
}
size_t scl_ws_bytes(int Bv, int T, int D) {
  (void)D;
  return scl_ws_layout(Bv * 2 * T, nullptr, nullptr);
}

// ---- prep: M = sum(mask), ordered index lists of valid / masked rows, zeroed accumulators --------------------
// Single CTA, block-wide scan per 1024-row chunk: deterministic order, no host round trip.
__global__ void scl_prep_kernel(const float* __restrict__ masks, int N, SclWs w, float* loss_out) {
  pdl_entry();
  // single block, 1024 threads, chunked scan
  __shared__ int scan[1024];
  __shared__ int base_v, base_m;
  __shared__ float red[32];
  if (threadIdx.x == 0) { base_v = 0; base_m = 0; }
  float s = 0.f;
  __syncthreads();
  for (int i0 = 0; i0 < N; i0 += 1024) {
    int i = i0 + threadIdx.x;
    int v = (i < N && masks[i] != 0.f) ? 1 : 0;
    if (i < N) { s += masks[i]; w.zext[i] = 0.f; w.c[i] = 0.f; }
    scan[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = threadIdx.x >= o ? scan[threadIdx.x - o] : 0;
      __syncthreads();
      scan[threadIdx.x] += t;
      __syncthreads();
    }
    int incl = scan[threadIdx.x];
    int total = scan[1023];
    int nin = (N - i0 < 1024) ? N - i0 : 1024;
    if (i < N) {
      if (v) w.valid[base_v + incl - 1] = i;
      else w.masked[base_m + (threadIdx.x + 1 - incl) - 1] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) { base_v += total; base_m += nin - total; }
    __syncthreads();
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 32; ++k) t += red[k];
    *w.M = t;
    *loss_out = 0.f;
    w.counts[0] = base_v;
    w.counts[1] = base_m;
  }
}

// ---- prep for large batches (N > 8192): the same outputs from three launches -----------------------------------------
// (1) per-chunk counts, M (mask sum: integers, exact in fp32 whatever the order), zeroed accumulators;
// (2) one CTA scans the chunk counts; (3) every chunk scans locally and writes its slice of the ordered lists.
__global__ void __launch_bounds__(1024) scl_prep_count_kernel(const float* __restrict__ masks, int N, SclWs w, float* loss_out) {
  pdl_entry();
  __shared__ int cnt[32];
  __shared__ float sum[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const float m = i < N ? masks[i] : 0.f;
  if (i < N) { w.zext[i] = 0.f; w.c[i] = 0.f; }
  const int v = __popc(__ballot_sync(0xffffffffu, i < N && m != 0.f));
  const float ms = warp_sum(m);
  if ((threadIdx.x & 31) == 0) { cnt[threadIdx.x >> 5] = v; sum[threadIdx.x >> 5] = ms; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    float t = 0.f;
    for (int k = 0; k < 32; ++k) { c += cnt[k]; t += sum[k]; }
    w.chunk[blockIdx.x] = c;
    if (t != 0.f) atomicAdd(w.M, t);
    if (blockIdx.x == 0) *loss_out = 0.f;
  }
}
__global__ void __launch_bounds__(1024) scl_prep_scan_kernel(int N, int nchunks, SclWs w) {
  pdl_entry();
  // exclusive scan of chunk[0..nchunks) -> chunk[nchunks+1 ..] (valid bases); masked base = 1024*c - valid base
  __shared__ int scan[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int* base = w.chunk + nchunks + 1;
  for (int c0 = 0; c0 < nchunks; c0 += 1024) {
    const int c = c0 + threadIdx.x;
    const int v = c < nchunks ? w.chunk[c] : 0;
    scan[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? scan[threadIdx.x - o] : 0;
      __syncthreads();
      scan[threadIdx.x] += t;
      __syncthreads();
    }
    if (c < nchunks) base[c] = carry + scan[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += scan[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    w.counts[0] = carry;
    w.counts[1] = N - carry;
  }
}
__global__ void __launch_bounds__(1024) scl_prep_write_kernel(const float* __restrict__ masks, int N, int nchunks, SclWs w) {
  pdl_entry();
  __shared__ int scan[1024];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int v = (i < N && masks[i] != 0.f) ? 1 : 0;
  scan[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int t = threadIdx.x >= o ? scan[threadIdx.x - o] : 0;
    __syncthreads();
    scan[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < N) {
    const int bv = w.chunk[nchunks + 1 + blockIdx.x];
    const int bm = blockIdx.x * 1024 - bv;
    const int incl = scan[threadIdx.x];
    if (v) w.valid[bv + incl - 1] = i;
    else w.masked[bm + (threadIdx.x + 1 - incl) - 1] = i;
  }
}

// 4-way unrolled dot product of two shared-memory vectors (16-byte aligned, D % 4 == 0).  Row i vs row j and row j vs
// row i go through the same sequence of operations, so l_ij == l_ji bit for bit.
__device__ __forceinline__ float dot4(const float* __restrict__ a, const float* __restrict__ b, int D) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
  for (int d = 0; d < D; d += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a + d);
    const float4 y = *reinterpret_cast<const float4*>(b + d);
    s0 = fmaf(x.x, y.x, s0);
    s1 = fmaf(x.y, y.y, s1);
    s2 = fmaf(x.z, y.z, s2);
    s3 = fmaf(x.w, y.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// cooperative, 16-byte vectorised load of up to 32 embedding rows into a padded shared-memory tile
__device__ __forceinline__ void load_tile(float* tile, int Dp, const float* __restrict__ embs, int D, const int* idx_list,
                                          int first, int count_limit, int base_row) {
  const int D4 = D >> 2;
  for (int i = threadIdx.x; i < 32 * D4; i += blockDim.x) {
    const int jj = i / D4, d4 = i - jj * D4;
    const int cs = first + jj;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cs < count_limit) {
      const int64_t row = idx_list ? (int64_t)idx_list[cs] : (int64_t)(base_row + cs);
      v = *reinterpret_cast<const float4*>(embs + row * D + 4 * d4);
    }
    *reinterpret_cast<float4*>(tile + jj * Dp + 4 * d4) = v;
  }
}

// ---- generic cross pass -------------------------------------------------------------------------------------------
// rows r in row list, columns k in col list:  e = cc_k * [not excluded] * exp(<e_r, e_k>/tau)
//   sum_out[r] += sum_k e                        (if sum_out)
//   vec_out[r,:] += rc_r * sum_k e * e_k / tau   (if vec_out)
// exclusion: same video (vid = idx / (2T)) when excl_same_video.
__global__ void __launch_bounds__(256)
scl_cross_kernel(const float* __restrict__ embs, int D, int T2, float inv_tau_div, const int* __restrict__ row_idx,
                 const int* __restrict__ row_cnt, const int* __restrict__ col_idx, const int* __restrict__ col_cnt,
                 const float* __restrict__ rc_arr, float rc_const, const float* __restrict__ cc_arr, float cc_const,
                 int excl_same_video, float* __restrict__ sum_out, float* __restrict__ vec_out) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  const int Dp = D + 4;
  float* tile = sm;                 // [32][Dp]
  float* er = tile + 32 * Dp;       // [8][D]
  float* ccs = er + 8 * D;          // [32]
  int* cvid = (int*)(ccs + 32);     // [32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrows = *row_cnt, ncols = *col_cnt;
  if (nrows == 0 || ncols == 0) return;   // nothing to couple (e.g. no masked frame in the batch): uniform exit
  for (int rb = blockIdx.x; rb * 8 < nrows; rb += gridDim.x) {   // row blocks, grid-stride (the grid is sized for the SMs)
    const int rslot = rb * 8 + warp;
    const bool ractive = rslot < nrows;
    const int r = ractive ? row_idx[rslot] : 0;
    const int rvid = r / T2;
    __syncthreads();   // the previous row block's readers of er are done
    for (int d = lane; d < D; d += 32) er[warp * D + d] = ractive ? embs[(int64_t)r * D + d] : 0.f;
    float acc[SCL_MAXD / 32];
#pragma unroll
    for (int k = 0; k < SCL_MAXD / 32; ++k) acc[k] = 0.f;
    float rsum = 0.f;
    // columns are split over blockIdx.y (outputs are accumulated with atomics)
    for (int c0 = blockIdx.y * 32; c0 < ncols; c0 += 32 * gridDim.y) {
      __syncthreads();
      load_tile(tile, Dp, embs, D, col_idx, c0, ncols, 0);
      if (threadIdx.x < 32) {
        const int cs = c0 + threadIdx.x;
        if (cs < ncols) {
          const int k = col_idx[cs];
          ccs[threadIdx.x] = cc_const * (cc_arr ? cc_arr[k] : 1.f);
          cvid[threadIdx.x] = k / T2;
        } else {
          ccs[threadIdx.x] = 0.f;
          cvid[threadIdx.x] = -1;
        }
      }
      __syncthreads();
      const float dot = dot4(er + warp * D, tile + lane * Dp, D);
      float wgt = ccs[lane];
      if (excl_same_video && cvid[lane] == rvid) wgt = 0.f;
      const float ex = (wgt != 0.f && ractive) ? wgt * expf(__fdiv_rn(dot, inv_tau_div)) : 0.f;
      rsum += ex;
      if (vec_out) {
        for (int jj = 0; jj < 32; ++jj) {
          const float gx = __shfl_sync(0xffffffffu, ex, jj);
          if (gx != 0.f) {
#pragma unroll
            for (int k = 0; k < SCL_MAXD / 32; ++k) {
              const int d = lane + 32 * k;
              if (d < D) acc[k] = fmaf(gx, tile[jj * Dp + d], acc[k]);
            }
          }
        }
      }
    }
    rsum = warp_sum(rsum);
    if (ractive) {
      if (sum_out && lane == 0 && rsum != 0.f) atomicAdd(sum_out + r, rsum);
      if (vec_out) {
        const float rc = rc_const * (rc_arr ? rc_arr[r] : 1.f);
#pragma unroll
        for (int k = 0; k < SCL_MAXD / 32; ++k) {
          const int d = lane + 32 * k;
          if (d < D && acc[k] != 0.f) atomicAdd(vec_out + (int64_t)r * D + d, __fdiv_rn(rc * acc[k], inv_tau_div));
        }
      }
    }
  }
}

// ---- masked-column passes with the exp matrix kept (scl.py:80 quirk; small batches) -----------------------------------
// The three generic cross passes above recompute the (valid row, masked column) dot products and exponentials three
// times and accumulate with atomics.  For N <= SCL_MEX_MAXN the matrix x_rk = 1e-6 e^{l_rk} is computed ONCE, stored
// ([n_valid x n_masked], a few hundred KB at the named shapes) and reused:
//   scl_mex_kernel    x_rk and its row sums (-> zext), 8 valid rows per CTA, no column split, no atomics on the matrix
//   scl_mgrad_kernel  after the pair kernel has produced c_r:   role A (valid rows, warp per row)  dE_r += c_r sum_k x_rk e_k / tau
//                                                               role B (masked rows, CTA per row)  dE_k += sum_r c_r x_rk e_r / tau
// Every output row is owned by one warp / CTA, so the updates are plain read-modify-writes.
__global__ void __launch_bounds__(256)
scl_mex_kernel(const float* __restrict__ embs, int D, float tau, SclWs w) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  const int Dp = D + 4;
  float* tile = sm;                 // [32][Dp]
  float* er = tile + 32 * Dp;       // [8][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = w.counts[0], nm = w.counts[1];
  if (nv == 0 || nm == 0) return;
  for (int rb = blockIdx.x; rb * 8 < nv; rb += gridDim.x) {
    const int rslot = rb * 8 + warp;
    const bool ractive = rslot < nv;
    const int r = ractive ? w.valid[rslot] : 0;
    __syncthreads();
    for (int d = lane; d < D; d += 32) er[warp * D + d] = ractive ? embs[(int64_t)r * D + d] : 0.f;
    float rsum = 0.f;
    for (int c0 = blockIdx.y * 32; c0 < nm; c0 += 32 * gridDim.y) {   // column tiles are split over blockIdx.y
      __syncthreads();
      load_tile(tile, Dp, embs, D, w.masked, c0, nm, 0);
      __syncthreads();
      const float dot = dot4(er + warp * D, tile + lane * Dp, D);
      if (ractive && c0 + lane < nm) {
        const float x = 1e-6f * expf(__fdiv_rn(dot, tau));
        w.mex[(size_t)rslot * nm + c0 + lane] = x;
        rsum += x;
      }
    }
    rsum = warp_sum(rsum);
    if (ractive && lane == 0 && rsum != 0.f) atomicAdd(w.zext + r, rsum);
  }
}

// One CTA = 8 output rows (a warp each) x one share (blockIdx.y) of the 32-row tiles of the other side, staged in shared
// memory.  role A (blockIdx.x < gridA): outputs = valid rows r, other side = masked rows k, weight c_r x_rk;
// role B: outputs = masked rows k, other side = valid rows r, same weight.  dE accumulates with one atomicAdd per element
// and share.
__global__ void __launch_bounds__(256)
scl_mgrad_kernel(const float* __restrict__ embs, int D, float tau, SclWs w, int gridA, float* __restrict__ d_embs) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  const int Dp = D + 4;
  float* tile = sm;                 // [32][Dp]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = w.counts[0], nm = w.counts[1];
  if (nv == 0 || nm == 0) return;
  const bool roleA = (int)blockIdx.x < gridA;
  const int nout = roleA ? nv : nm, nin = roleA ? nm : nv;
  const int* out_idx = roleA ? w.valid : w.masked;
  const int* in_idx = roleA ? w.masked : w.valid;
  const int gx = roleA ? gridA : (int)gridDim.x - gridA;
  const int bx = roleA ? (int)blockIdx.x : (int)blockIdx.x - gridA;
  const int D4 = D >> 2;
  for (int ob = bx; ob * 8 < nout; ob += gx) {
    const int oslot = ob * 8 + warp;
    const bool oactive = oslot < nout;
    const int orow = oactive ? out_idx[oslot] : 0;
    const float co = (roleA && oactive) ? w.c[orow] : 1.f;
    float4 acc[SCL_MAXD / 128];
#pragma unroll
    for (int u = 0; u < SCL_MAXD / 128; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i0 = blockIdx.y * 32; i0 < nin; i0 += 32 * gridDim.y) {
      __syncthreads();
      load_tile(tile, Dp, embs, D, in_idx, i0, nin, 0);
      // this lane's weight for inner row i0 + lane
      float x = 0.f;
      if (oactive && i0 + lane < nin) {
        if (roleA) x = co * w.mex[(size_t)oslot * nm + i0 + lane];
        else x = w.c[in_idx[i0 + lane]] * w.mex[(size_t)(i0 + lane) * nm + oslot];
      }
      __syncthreads();
      if (__any_sync(0xffffffffu, x != 0.f)) {
        for (int jj = 0; jj < 32; ++jj) {
          const float xj = __shfl_sync(0xffffffffu, x, jj);
          if (xj != 0.f) {
#pragma unroll
            for (int u = 0; u < SCL_MAXD / 128; ++u) {
              const int d4 = lane + 32 * u;
              if (d4 < D4) {
                const float4 e = *reinterpret_cast<const float4*>(tile + jj * Dp + 4 * d4);
                acc[u].x = fmaf(xj, e.x, acc[u].x); acc[u].y = fmaf(xj, e.y, acc[u].y);
                acc[u].z = fmaf(xj, e.z, acc[u].z); acc[u].w = fmaf(xj, e.w, acc[u].w);
              }
            }
          }
        }
      }
    }
    if (oactive) {
      float* out = d_embs + (int64_t)orow * D;
#pragma unroll
      for (int u = 0; u < SCL_MAXD / 128; ++u) {
        const int d4 = lane + 32 * u;
        if (d4 < D4) {
          if (acc[u].x != 0.f) atomicAdd(out + 4 * d4, __fdiv_rn(acc[u].x, tau));
          if (acc[u].y != 0.f) atomicAdd(out + 4 * d4 + 1, __fdiv_rn(acc[u].y, tau));
          if (acc[u].z != 0.f) atomicAdd(out + 4 * d4 + 2, __fdiv_rn(acc[u].z, tau));
          if (acc[u].w != 0.f) atomicAdd(out + 4 * d4 + 3, __fdiv_rn(acc[u].w, tau));
        }
      }
    }
  }
}

// ---- per-row quantities against the partner block -------------------------------------------------------------------
__device__ __forceinline__ float ts_dist(float si, float Li, float Lj, float sj) {
  // torch: abs(steps_i / L_i * L_j - steps_j), all float32 ops (scl.py:62)
  return fabsf(__fsub_rn(__fmul_rn(__fdiv_rn(si, Li), Lj), sj));
}

// One warp per row; lanes own partner columns j = lane + 32*k.  PHASE 0: row stats + loss.  PHASE 1: gradient.
template <int PHASE>
__global__ void __launch_bounds__(256)
scl_pair_kernel(const float* __restrict__ embs, const int64_t* __restrict__ seq_lens, const int64_t* __restrict__ steps,
                const float* __restrict__ masks, int T, int D, float tau, float two_var, SclWs w,
                float* __restrict__ loss_out, float* __restrict__ d_embs) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  const int Dp = D + 4;
  float* tile = sm;            // [32][Dp] partner embeddings of the current column chunk
  float* er = tile + 32 * Dp;  // [8][D]   row embeddings
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blocks_per_view = (T + 7) / 8;
  const int vv = blockIdx.x / blocks_per_view;        // (video, view) index
  const int t = (blockIdx.x % blocks_per_view) * 8 + warp;
  const int v = vv >> 1, a = vv & 1;
  const bool ractive = t < T;
  const int i = vv * T + (ractive ? t : 0);           // global row
  const int pbase = (v * 2 + (1 - a)) * T;             // first row of the partner block
  const float M = *w.M;
  const float Li = (float)seq_lens[v * 2 + a], Lj = (float)seq_lens[v * 2 + (1 - a)];
  const float mi = ractive ? masks[i] : 0.f;
  const float si = (float)steps[i];
  for (int d = lane; d < D; d += 32) er[warp * D + d] = embs[(int64_t)i * D + d];

  float l[SCL_MAXTC], pw[SCL_MAXTC];
  const int nch = (T + 31) / 32;
  // pass over column chunks: logits and Gaussian label weights into registers
#pragma unroll
  for (int k = 0; k < SCL_MAXTC; ++k) {
    l[k] = 0.f;
    pw[k] = 0.f;
    if (k < nch) {
      __syncthreads();
      load_tile(tile, Dp, embs, D, nullptr, k * 32, T, pbase);
      __syncthreads();
      const int j = k * 32 + lane;
      const float dot = dot4(er + warp * D, tile + lane * Dp, D);
      if (j < T) {
        l[k] = __fdiv_rn(dot, tau);
        const float mj = masks[pbase + j];
        if (mi != 0.f && mj != 0.f) {
          const float dd = ts_dist(si, Li, Lj, (float)steps[pbase + j]);
          pw[k] = expf(__fdiv_rn(-(dd * dd), two_var));
        }
      }
    }
  }
  // row reductions (valid entries are exactly those with mm = 1; pw is 0 elsewhere but e^l needs the mask)
  float den = 0.f, zp = 0.f;
#pragma unroll
  for (int k = 0; k < SCL_MAXTC; ++k) {
    if (k < nch) {
      const int j = k * 32 + lane;
      const bool mm = j < T && mi != 0.f && masks[pbase + j] != 0.f;
      den += pw[k];
      if (mm) zp += expf(l[k]);
    }
  }
  den = warp_sum(den);
  zp = warp_sum(zp);

  if (PHASE == 0) {
    const float Z = zp + w.zext[i];
    float g = 0.f, loss = 0.f;
    if (mi != 0.f && Z > 0.f) {
#pragma unroll
      for (int k = 0; k < SCL_MAXTC; ++k) {
        if (k < nch) {
          const int j = k * 32 + lane;
          const bool mm = j < T && masks[pbase + j] != 0.f;
          if (mm) {
            const float y = den > 0.f ? __fdiv_rn(pw[k], den) : 0.f;
            const float p = __fdiv_rn(expf(l[k]), Z);
            const float q = p + 1e-6f;
            if (y > 0.f) {
              loss += y * (logf(y) - logf(q));
              g += y * __fdiv_rn(p, q);
            }
          }
        }
      }
    }
    g = warp_sum(g);
    loss = warp_sum(loss);
    if (ractive && lane == 0) {
      w.Z[i] = Z;
      w.g[i] = g;
      w.den[i] = den;
      w.c[i] = (mi != 0.f && Z > 0.f) ? g / (Z * M) : 0.f;
      if (loss != 0.f) atomicAdd(loss_out, loss / M);
    }
    return;
  }

  // PHASE 1: dE_i = sum_j (G_ij + G_ji) e_j / tau over the partner block
  const float Zi = w.Z[i], gi = w.g[i];
  float acc[SCL_MAXD / 32];
#pragma unroll
  for (int k = 0; k < SCL_MAXD / 32; ++k) acc[k] = 0.f;
#pragma unroll
  for (int k = 0; k < SCL_MAXTC; ++k) {
    if (k < nch) {
      const int j = k * 32 + lane;
      float coef = 0.f;
      if (j < T && mi != 0.f && masks[pbase + j] != 0.f) {
        const float ex = expf(l[k]);
        // own direction (row i)
        if (Zi > 0.f) {
          const float y = den > 0.f ? __fdiv_rn(pw[k], den) : 0.f;
          const float p = __fdiv_rn(ex, Zi);
          const float r = __fdiv_rn(p, p + 1e-6f);
          coef += p * gi - y * r;
        }
        // partner direction (row j, column i): its own float32 timestamp rounding and its own statistics
        const int jr = pbase + j;
        const float Zj = w.Z[jr];
        if (Zj > 0.f) {
          const float ddj = ts_dist((float)steps[jr], Lj, Li, si);
          const float pwj = expf(__fdiv_rn(-(ddj * ddj), two_var));
          const float denj = w.den[jr];
          const float yj = denj > 0.f ? __fdiv_rn(pwj, denj) : 0.f;
          const float pj = __fdiv_rn(ex, Zj);
          const float rj = __fdiv_rn(pj, pj + 1e-6f);
          coef += pj * w.g[jr] - yj * rj;
        }
        coef = coef / M;
      }
      // reload the partner chunk and accumulate coef * e_j over lanes' d-slices
      __syncthreads();
      load_tile(tile, Dp, embs, D, nullptr, k * 32, T, pbase);
      __syncthreads();
      for (int jj = 0; jj < 32; ++jj) {
        const float cf = __shfl_sync(0xffffffffu, coef, jj);
        if (cf != 0.f) {
#pragma unroll
          for (int kk = 0; kk < SCL_MAXD / 32; ++kk) {
            const int d = lane + 32 * kk;
            if (d < D) acc[kk] = fmaf(cf, tile[jj * Dp + d], acc[kk]);
          }
        }
      }
    }
  }
  if (ractive) {
#pragma unroll
    for (int kk = 0; kk < SCL_MAXD / 32; ++kk) {
      const int d = lane + 32 * kk;
      if (d < D) d_embs[(int64_t)i * D + d] = __fdiv_rn(acc[kk], tau);
    }
  }
}

// =====================================================================================================================
// Fused per-pair kernel: ONE CTA per video pair does the whole own-pair part of the loss and of its gradient.
//   stage E0, E1 [T, D] in shared memory (the only HBM reads: 2*T*D*4 B + steps/masks), ex = exp(E0 E1^T / tau) [T, T],
//   row statistics of both view directions (warp per row, shuffle reductions), the coefficient matrix
//   coef_ij = dloss/dl_ij of both directions, dE0 = coef E1 / tau, dE1 = coef^T E0 / tau (the only HBM writes).
// The T x T block never leaves shared memory and every embedding row is read from HBM exactly once, so the kernel is
// bound by 4*T*D*4 bytes per pair.  zext (masked-column / batch-negative extras of Z, produced by the cross pass before
// this kernel) is read per row; c_i = g_i / (Z_i M) is written for the gradient cross passes that follow.
// =====================================================================================================================
constexpr int SCL_RG = 10;   // rows per register tile of the dE phase (RG float4 accumulators per thread)

template <int NT>
__global__ void __launch_bounds__(NT)
scl_pair_fused_kernel(const float* __restrict__ embs, const int64_t* __restrict__ seq_lens, const int64_t* __restrict__ steps,
                      const float* __restrict__ masks, int T, int D, float tau, float two_var, SclWs w,
                      float* __restrict__ loss_out, float* __restrict__ d_embs) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  const int Dp = D + 4, Tp = T + 1;
  float* E = sm;                          // [2T][Dp]  rows 0..T-1 = view 0, T..2T-1 = view 1
  float* EX = E + (size_t)2 * T * Dp;     // [T][Tp]   exp(l_ij), later coef_ij   (i: view 0, j: view 1)
  float* st = EX + (size_t)T * Tp;        // [2T] steps as float
  float* mk = st + 2 * T;                 // [2T] masks
  float* Zs = mk + 2 * T;                 // [2T] partition sums
  float* gs = Zs + 2 * T;                 // [2T] g
  float* dn = gs + 2 * T;                 // [2T] label normalisers
  __shared__ float red_loss[NT / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v = blockIdx.x;
  const int64_t row0 = (int64_t)v * 2 * T;
  const float M = *w.M;
  const float L0 = (float)seq_lens[v * 2], L1 = (float)seq_lens[v * 2 + 1];

  // ---- stage the pair ----
  const int D4 = D >> 2;
  for (int i = tid; i < 2 * T * D4; i += NT) {
    const int r = i / D4, d4 = i - r * D4;
    *reinterpret_cast<float4*>(E + r * Dp + 4 * d4) = *reinterpret_cast<const float4*>(embs + (row0 + r) * D + 4 * d4);
  }
  for (int i = tid; i < 2 * T; i += NT) {
    st[i] = (float)steps[row0 + i];
    mk[i] = masks[row0 + i];
  }
  __syncthreads();

  // ---- ex_ij = exp(<e0_i, e1_j> / tau): 2 x 2 register tiles ----
  const int T2h = (T + 1) >> 1;
  for (int tix = tid; tix < T2h * T2h; tix += NT) {
    const int ti = tix / T2h, tj = tix - ti * T2h;
    const int i0 = 2 * ti, i1 = min(2 * ti + 1, T - 1), j0 = 2 * tj, j1 = min(2 * tj + 1, T - 1);
    const float* a0 = E + i0 * Dp;
    const float* a1 = E + i1 * Dp;
    const float* b0 = E + (T + j0) * Dp;
    const float* b1 = E + (T + j1) * Dp;
    float s00[4] = {0.f, 0.f, 0.f, 0.f}, s01[4] = {0.f, 0.f, 0.f, 0.f}, s10[4] = {0.f, 0.f, 0.f, 0.f}, s11[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int d = 0; d < D; d += 4) {
      const float4 x0 = *reinterpret_cast<const float4*>(a0 + d), x1 = *reinterpret_cast<const float4*>(a1 + d);
      const float4 y0 = *reinterpret_cast<const float4*>(b0 + d), y1 = *reinterpret_cast<const float4*>(b1 + d);
      s00[0] = fmaf(x0.x, y0.x, s00[0]); s00[1] = fmaf(x0.y, y0.y, s00[1]); s00[2] = fmaf(x0.z, y0.z, s00[2]); s00[3] = fmaf(x0.w, y0.w, s00[3]);
      s01[0] = fmaf(x0.x, y1.x, s01[0]); s01[1] = fmaf(x0.y, y1.y, s01[1]); s01[2] = fmaf(x0.z, y1.z, s01[2]); s01[3] = fmaf(x0.w, y1.w, s01[3]);
      s10[0] = fmaf(x1.x, y0.x, s10[0]); s10[1] = fmaf(x1.y, y0.y, s10[1]); s10[2] = fmaf(x1.z, y0.z, s10[2]); s10[3] = fmaf(x1.w, y0.w, s10[3]);
      s11[0] = fmaf(x1.x, y1.x, s11[0]); s11[1] = fmaf(x1.y, y1.y, s11[1]); s11[2] = fmaf(x1.z, y1.z, s11[2]); s11[3] = fmaf(x1.w, y1.w, s11[3]);
    }
    // same summation tree as dot4() of the row-warp kernels
    EX[i0 * Tp + j0] = expf(__fdiv_rn((s00[0] + s00[1]) + (s00[2] + s00[3]), tau));
    EX[i0 * Tp + j1] = expf(__fdiv_rn((s01[0] + s01[1]) + (s01[2] + s01[3]), tau));
    EX[i1 * Tp + j0] = expf(__fdiv_rn((s10[0] + s10[1]) + (s10[2] + s10[3]), tau));
    EX[i1 * Tp + j1] = expf(__fdiv_rn((s11[0] + s11[1]) + (s11[2] + s11[3]), tau));
  }
  __syncthreads();

  // ---- row statistics of both directions: warp per row r (r < T: view-0 row i over j; r >= T: view-1 row j over i) ----
  float loss_acc = 0.f;
  for (int r = warp; r < 2 * T; r += NT / 32) {
    const bool dir1 = r >= T;
    const int a = dir1 ? r - T : r;                 // index inside the own view
    const float mi = mk[r], si = st[r];
    const float Li = dir1 ? L1 : L0, Lj = dir1 ? L0 : L1;
    const int pb = dir1 ? 0 : T;                    // partner block offset in st / mk
    float den = 0.f, zp = 0.f;
    for (int b = lane; b < T; b += 32) {
      const bool mm = mi != 0.f && mk[pb + b] != 0.f;
      if (mm) {
        const float dd = ts_dist(si, Li, Lj, st[pb + b]);
        den += expf(__fdiv_rn(-(dd * dd), two_var));
        zp += dir1 ? EX[b * Tp + a] : EX[a * Tp + b];
      }
    }
    den = warp_sum(den);
    zp = warp_sum(zp);
    const float Z = zp + w.zext[row0 + r];
    float g = 0.f, loss = 0.f;
    if (mi != 0.f && Z > 0.f) {
      for (int b = lane; b < T; b += 32) {
        if (mk[pb + b] != 0.f) {
          const float dd = ts_dist(si, Li, Lj, st[pb + b]);
          const float pw = expf(__fdiv_rn(-(dd * dd), two_var));
          const float y = den > 0.f ? __fdiv_rn(pw, den) : 0.f;
          const float p = __fdiv_rn(dir1 ? EX[b * Tp + a] : EX[a * Tp + b], Z);
          const float q = p + 1e-6f;
          if (y > 0.f) {
            loss += y * (logf(y) - logf(q));
            g += y * __fdiv_rn(p, q);
          }
        }
      }
    }
    g = warp_sum(g);
    loss = warp_sum(loss);
    if (lane == 0) {
      Zs[r] = Z;
      gs[r] = g;
      dn[r] = den;
      w.c[row0 + r] = (mi != 0.f && Z > 0.f) ? g / (Z * M) : 0.f;
      loss_acc += loss;
    }
  }
  if (lane == 0) red_loss[warp] = loss_acc;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int k = 0; k < NT / 32; ++k) t += red_loss[k];
    if (t != 0.f) atomicAdd(loss_out, t / M);
  }
  if (d_embs == nullptr) return;

  // ---- coef_ij = ( [p g_i - y r]_dir0 + [p' g'_j - y' r']_dir1 ) / M, in place over ex ----
  for (int ix = tid; ix < T * T; ix += NT) {
    const int i = ix / T, j = ix - i * T;
    float coef = 0.f;
    if (mk[i] != 0.f && mk[T + j] != 0.f) {
      const float ex = EX[i * Tp + j];
      const float Zi = Zs[i], Zj = Zs[T + j];
      if (Zi > 0.f) {
        const float dd = ts_dist(st[i], L0, L1, st[T + j]);
        const float pw = expf(__fdiv_rn(-(dd * dd), two_var));
        const float y = dn[i] > 0.f ? __fdiv_rn(pw, dn[i]) : 0.f;
        const float p = __fdiv_rn(ex, Zi);
        coef += p * gs[i] - y * __fdiv_rn(p, p + 1e-6f);
      }
      if (Zj > 0.f) {
        const float dd = ts_dist(st[T + j], L1, L0, st[i]);
        const float pw = expf(__fdiv_rn(-(dd * dd), two_var));
        const float y = dn[T + j] > 0.f ? __fdiv_rn(pw, dn[T + j]) : 0.f;
        const float p = __fdiv_rn(ex, Zj);
        coef += p * gs[T + j] - y * __fdiv_rn(p, p + 1e-6f);
      }
      coef = coef / M;
    }
    EX[i * Tp + j] = coef;
  }
  __syncthreads();

  // ---- dE0_i = sum_j coef_ij e1_j / tau ; dE1_j = sum_i coef_ij e0_i / tau : register tiles of SCL_RG rows x 4 channels ----
  const int ngroups = (T + SCL_RG - 1) / SCL_RG;
  const int ntasks = 2 * ngroups * D4;               // (view, row group, 4-channel chunk)
  for (int task = tid; task < ntasks; task += NT) {
    const int d4 = task % D4;
    const int gr = (task / D4) % ngroups;
    const int view = task / (D4 * ngroups);
    const int r0 = gr * SCL_RG;
    const float* other = E + (view ? 0 : T) * Dp + 4 * d4;   // partner rows
    float4 acc[SCL_RG];
#pragma unroll
    for (int k = 0; k < SCL_RG; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < T; ++b) {
      const float4 x = *reinterpret_cast<const float4*>(other + b * Dp);
#pragma unroll
      for (int k = 0; k < SCL_RG; ++k) {
        const int a = min(r0 + k, T - 1);
        const float cf = view ? EX[b * Tp + a] : EX[a * Tp + b];
        acc[k].x = fmaf(cf, x.x, acc[k].x); acc[k].y = fmaf(cf, x.y, acc[k].y);
        acc[k].z = fmaf(cf, x.z, acc[k].z); acc[k].w = fmaf(cf, x.w, acc[k].w);
      }
    }
#pragma unroll
    for (int k = 0; k < SCL_RG; ++k) {
      const int a = r0 + k;
      if (a < T)
        *reinterpret_cast<float4*>(d_embs + (row0 + view * T + a) * D + 4 * d4) =
            make_float4(__fdiv_rn(acc[k].x, tau), __fdiv_rn(acc[k].y, tau), __fdiv_rn(acc[k].z, tau), __fdiv_rn(acc[k].w, tau));
    }
  }
}

// ---- second version of the fused pair kernel: the same algorithm with the per-element work cut down --------------------
//  * the exponents of the Gaussian label weights of both directions are computed once, next to exp(l_ij), and kept in
//    shared memory (the first version re-evaluated ts_dist + expf in the statistics pass and again in the coefficient
//    pass); labels are formed as y = 2^(a - log2 den): one MUFU, no division, safe when den is denormal;
//  * s_r / L_r (one IEEE division per row, exactly torch's float32 op) is hoisted out of the T x T loops;
//  * per-row reciprocals replace per-element IEEE divisions, exp2f / __logf / __fdividef replace expf / logf / division
//    where the operand is not a timestamp (relative error <= 1e-6 on p, y, r; the 1e-5 parity tests stay green).
template <int NT>
__global__ void __launch_bounds__(NT)
scl_pair_fused2_kernel(const float* __restrict__ embs, const int64_t* __restrict__ seq_lens, const int64_t* __restrict__ steps,
                       const float* __restrict__ masks, int T, int D, float tau, float two_var, SclWs w,
                       float* __restrict__ loss_out, float* __restrict__ d_embs) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];
  const int Dp = D + 4, Tp = T + 1;
  float* E = sm;                          // [2T][Dp]
  float* EX = E + (size_t)2 * T * Dp;     // [T][Tp]  exp(l_ij), later coef_ij          (i: view 0, j: view 1)
  float* PW0 = EX + (size_t)T * Tp;       // [T][Tp]  log2 of the label weight of direction 0 at (i, j); -inf = masked pair
  float* PW1 = PW0 + (size_t)T * Tp;      // [T][Tp]  same for direction 1 at (j, i), stored at [i][j]
  float* st = PW1 + (size_t)T * Tp;       // [2T] steps as float
  float* mk = st + 2 * T;                 // [2T] masks
  float* ar = mk + 2 * T;                 // [2T] fl(s_r / L_r)
  float* iZ = ar + 2 * T;                 // [2T] 1 / Z   (0 when the row is masked or Z = 0)
  float* gs = iZ + 2 * T;                 // [2T] g
  float* iD = gs + 2 * T;                 // [2T] log2(den) of the label normaliser (+inf when den = 0): y = 2^(a - log2 den),
                                          //      no division and no overflow when every partner is far (denormal den)
  __shared__ float red_loss[NT / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v = blockIdx.x;
  const int64_t row0 = (int64_t)v * 2 * T;
  const float M = *w.M;
  const float invM = 1.f / M, inv_tau = 1.f / tau;
  const float L0 = (float)seq_lens[v * 2], L1 = (float)seq_lens[v * 2 + 1];
  const float c_ex = 1.4426950408889634f / tau;           // exp(x / tau) = 2^(x * c_ex)
  const float c_pw = -1.4426950408889634f / two_var;      // exp(-d^2 / (2 var)) = 2^(d^2 * c_pw)

  const int D4 = D >> 2;
  for (int i = tid; i < 2 * T * D4; i += NT) {
    const int r = i / D4, d4 = i - r * D4;
    *reinterpret_cast<float4*>(E + r * Dp + 4 * d4) = *reinterpret_cast<const float4*>(embs + (row0 + r) * D + 4 * d4);
  }
  for (int i = tid; i < 2 * T; i += NT) {
    const float s = (float)steps[row0 + i];
    st[i] = s;
    mk[i] = masks[row0 + i];
    ar[i] = __fdiv_rn(s, i < T ? L0 : L1);
  }
  __syncthreads();

  // ---- ex_ij, pw0_ij, pw1_ij: 2 x 2 register tiles ----
  const int T2h = (T + 1) >> 1;
  for (int tix = tid; tix < T2h * T2h; tix += NT) {
    const int ti = tix / T2h, tj = tix - ti * T2h;
    const int ii[2] = {2 * ti, min(2 * ti + 1, T - 1)}, jj[2] = {2 * tj, min(2 * tj + 1, T - 1)};
    const float* a0 = E + ii[0] * Dp;
    const float* a1 = E + ii[1] * Dp;
    const float* b0 = E + (T + jj[0]) * Dp;
    const float* b1 = E + (T + jj[1]) * Dp;
    float s00[4] = {0.f, 0.f, 0.f, 0.f}, s01[4] = {0.f, 0.f, 0.f, 0.f}, s10[4] = {0.f, 0.f, 0.f, 0.f}, s11[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int d = 0; d < D; d += 4) {
      const float4 x0 = *reinterpret_cast<const float4*>(a0 + d), x1 = *reinterpret_cast<const float4*>(a1 + d);
      const float4 y0 = *reinterpret_cast<const float4*>(b0 + d), y1 = *reinterpret_cast<const float4*>(b1 + d);
      s00[0] = fmaf(x0.x, y0.x, s00[0]); s00[1] = fmaf(x0.y, y0.y, s00[1]); s00[2] = fmaf(x0.z, y0.z, s00[2]); s00[3] = fmaf(x0.w, y0.w, s00[3]);
      s01[0] = fmaf(x0.x, y1.x, s01[0]); s01[1] = fmaf(x0.y, y1.y, s01[1]); s01[2] = fmaf(x0.z, y1.z, s01[2]); s01[3] = fmaf(x0.w, y1.w, s01[3]);
      s10[0] = fmaf(x1.x, y0.x, s10[0]); s10[1] = fmaf(x1.y, y0.y, s10[1]); s10[2] = fmaf(x1.z, y0.z, s10[2]); s10[3] = fmaf(x1.w, y0.w, s10[3]);
      s11[0] = fmaf(x1.x, y1.x, s11[0]); s11[1] = fmaf(x1.y, y1.y, s11[1]); s11[2] = fmaf(x1.z, y1.z, s11[2]); s11[3] = fmaf(x1.w, y1.w, s11[3]);
    }
    const float dots[2][2] = {{(s00[0] + s00[1]) + (s00[2] + s00[3]), (s01[0] + s01[1]) + (s01[2] + s01[3])},
                              {(s10[0] + s10[1]) + (s10[2] + s10[3]), (s11[0] + s11[1]) + (s11[2] + s11[3])}};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int i = ii[a], j = jj[b];
        const bool mm = mk[i] != 0.f && mk[T + j] != 0.f;
        const float d0 = fabsf(__fsub_rn(__fmul_rn(ar[i], L1), st[T + j]));       // scl.py:62, direction 0 row i
        const float d1 = fabsf(__fsub_rn(__fmul_rn(ar[T + j], L0), st[i]));       // direction 1 row j
        EX[i * Tp + j] = exp2f(dots[a][b] * c_ex);
        PW0[i * Tp + j] = mm ? d0 * d0 * c_pw : -INFINITY;
        PW1[i * Tp + j] = mm ? d1 * d1 * c_pw : -INFINITY;
      }
    }
  }
  __syncthreads();

  // ---- row statistics of both directions: warp per row ----
  float loss_acc = 0.f;
  for (int r = warp; r < 2 * T; r += NT / 32) {
    const bool dir1 = r >= T;
    const int a = dir1 ? r - T : r;
    const float mi = mk[r];
    const int pb = dir1 ? 0 : T;
    const float* pwm = dir1 ? PW1 : PW0;
    float den = 0.f, zp = 0.f;
    for (int b = lane; b < T; b += 32) {
      const int ix = dir1 ? b * Tp + a : a * Tp + b;
      den += exp2f(pwm[ix]);
      if (mi != 0.f && mk[pb + b] != 0.f) zp += EX[ix];
    }
    den = warp_sum(den);
    zp = warp_sum(zp);
    const float Z = zp + w.zext[row0 + r];
    const bool live = mi != 0.f && Z > 0.f;
    const float invZ = live ? 1.f / Z : 0.f;
    const float l2den = den > 0.f ? log2f(den) : INFINITY;
    float g = 0.f, loss = 0.f;
    if (live) {
      for (int b = lane; b < T; b += 32) {
        const int ix = dir1 ? b * Tp + a : a * Tp + b;
        const float y = exp2f(pwm[ix] - l2den);
        // y > 0 implies both frames valid (pw is 0 for masked pairs).  Terms below 1e-30 are dropped: they change the loss
        // by < 1e-28, and __logf would flush a denormal y to zero and return -inf
        if (y > 1e-30f) {
          const float p = EX[ix] * invZ;
          const float q = p + 1e-6f;
          loss += y * (__logf(y) - __logf(q));
          g += y * __fdividef(p, q);
        }
      }
    }
    g = warp_sum(g);
    loss = warp_sum(loss);
    if (lane == 0) {
      iZ[r] = invZ;
      gs[r] = g;
      iD[r] = l2den;
      w.c[row0 + r] = live ? g * invZ * invM : 0.f;
      loss_acc += loss;
    }
  }
  if (lane == 0) red_loss[warp] = loss_acc;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int k = 0; k < NT / 32; ++k) t += red_loss[k];
    if (t != 0.f) atomicAdd(loss_out, t * invM);
  }
  if (d_embs == nullptr) return;

  // ---- coef_ij in place over ex ----
  for (int ix = tid; ix < T * T; ix += NT) {
    const int i = ix / T, j = ix - i * T;
    float coef = 0.f;
    if (mk[i] != 0.f && mk[T + j] != 0.f) {
      const float ex = EX[i * Tp + j];
      const float p0 = ex * iZ[i], p1 = ex * iZ[T + j];
      const float y0 = exp2f(PW0[i * Tp + j] - iD[i]), y1 = exp2f(PW1[i * Tp + j] - iD[T + j]);
      if (iZ[i] > 0.f) coef += p0 * gs[i] - y0 * __fdividef(p0, p0 + 1e-6f);
      if (iZ[T + j] > 0.f) coef += p1 * gs[T + j] - y1 * __fdividef(p1, p1 + 1e-6f);
      coef *= invM;
    }
    EX[i * Tp + j] = coef;
  }
  __syncthreads();

  // ---- dE0_i = sum_j coef_ij e1_j / tau ; dE1_j = sum_i coef_ij e0_i / tau ----
  const int ngroups = (T + SCL_RG - 1) / SCL_RG;
  const int ntasks = 2 * ngroups * D4;
  for (int task = tid; task < ntasks; task += NT) {
    const int d4 = task % D4;
    const int gr = (task / D4) % ngroups;
    const int view = task / (D4 * ngroups);
    const int r0 = gr * SCL_RG;
    const float* other = E + (view ? 0 : T) * Dp + 4 * d4;
    float4 acc[SCL_RG];
#pragma unroll
    for (int k = 0; k < SCL_RG; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < T; ++b) {
      const float4 x = *reinterpret_cast<const float4*>(other + b * Dp);
#pragma unroll
      for (int k = 0; k < SCL_RG; ++k) {
        const int a = min(r0 + k, T - 1);
        const float cf = view ? EX[b * Tp + a] : EX[a * Tp + b];
        acc[k].x = fmaf(cf, x.x, acc[k].x); acc[k].y = fmaf(cf, x.y, acc[k].y);
        acc[k].z = fmaf(cf, x.z, acc[k].z); acc[k].w = fmaf(cf, x.w, acc[k].w);
      }
    }
#pragma unroll
    for (int k = 0; k < SCL_RG; ++k) {
      const int a = r0 + k;
      if (a < T)
        *reinterpret_cast<float4*>(d_embs + (row0 + view * T + a) * D + 4 * d4) =
            make_float4(acc[k].x * inv_tau, acc[k].y * inv_tau, acc[k].z * inv_tau, acc[k].w * inv_tau);
    }
  }
}
static size_t scl_fused2_smem(int T, int D) {
  return ((size_t)2 * T * (D + 4) + (size_t)3 * T * (T + 1) + (size_t)12 * T) * sizeof(float);
}

static size_t scl_fused_smem(int T, int D) {
  return ((size_t)2 * T * (D + 4) + (size_t)T * (T + 1) + (size_t)10 * T) * sizeof(float);
}

int scl_fwd_bwd(const float* embs, const int64_t* seq_lens, const int64_t* steps, const float* masks, int Bv, int T,
                int D, float temperature, float label_variance, int negative_type, int quirk, float* loss_out,
                float* d_embs, void* ws, size_t ws_bytes, cudaStream_t st) {
  MVF_REQUIRE(embs && seq_lens && steps && masks && loss_out && ws, MVF_ERR_BAD_ARG, "scl: null pointer");
  MVF_REQUIRE(Bv > 0 && T > 0 && D > 0, MVF_ERR_BAD_ARG, "scl: bad shape Bv=%d T=%d D=%d", Bv, T, D);
  MVF_REQUIRE(D <= SCL_MAXD && D % 4 == 0, MVF_ERR_UNSUPPORTED, "scl: embedding size %d must be a multiple of 4 and <= %d", D,
              SCL_MAXD);
  MVF_REQUIRE((((uintptr_t)embs) & 15) == 0, MVF_ERR_ALIGN, "scl: embeddings must be 16-byte aligned");
  MVF_REQUIRE(T <= 32 * SCL_MAXTC, MVF_ERR_UNSUPPORTED, "scl: %d frames > %d", T, 32 * SCL_MAXTC);
  MVF_REQUIRE(negative_type == MVF_NEG_SINGLE_NOSELF || negative_type == MVF_NEG_BATCH_NOSELF, MVF_ERR_UNSUPPORTED,
              "scl: negative_type %d", negative_type);
  const int64_t N64 = (int64_t)Bv * 2 * T;
  MVF_REQUIRE(N64 < (1ll << 30), MVF_ERR_BAD_ARG, "scl: batch too large");
  const int N = (int)N64;
  SclWs w;
  size_t need = scl_ws_layout(N, &w, (char*)ws);
  MVF_REQUIRE(ws_bytes >= need, MVF_ERR_WORKSPACE, "scl: workspace %zu < %zu bytes", ws_bytes, need);

  if (N <= 8192) {
    launch_k(scl_prep_kernel, 1, 1024, 0, st, masks, N, w, loss_out);
    MVF_CHECK_LAUNCH();
  } else {
    const int nchunks = cdiv(N, 1024);
    MVF_CHECK_CUDA(cudaMemsetAsync(w.M, 0, sizeof(float), st));
    launch_k(scl_prep_count_kernel, nchunks, 1024, 0, st, masks, N, w, loss_out);
    MVF_CHECK_LAUNCH();
    launch_k(scl_prep_scan_kernel, 1, 1024, 0, st, N, nchunks, w);
    MVF_CHECK_LAUNCH();
    launch_k(scl_prep_write_kernel, nchunks, 1024, 0, st, masks, N, nchunks, w);
    MVF_CHECK_LAUNCH();
  }

  const size_t smem = ((size_t)32 * (D + 4) + 8 * D + 64) * sizeof(float);
  // cross passes: row blocks are visited grid-stride (row / column counts live on the device, so the grid is sized for the
  // machine, not for the worst case) and the column list is split over blockIdx.y (outputs accumulate with atomics)
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaDeviceProp prop;
    sms = (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) ? prop.multiProcessorCount : 148;
  }
  int cx = cdiv(N, 8);
  if (cx > sms * 2) cx = sms * 2;
  // (measured: fewer column splits for the valid-row passes make the named shapes slower -- 0.16 vs 0.12 ms at 32 pairs)
  const dim3 cross_grid(cx, 8), cross_grid_tall(cx, 8);
  const int T2 = 2 * T;
  const bool batch = negative_type == MVF_NEG_BATCH_NOSELF;
  // Z extras
  const bool use_mex = quirk && w.mex != nullptr;
  if (use_mex) {
    launch_k(scl_mex_kernel, dim3(cx, 8), 256, smem, st, embs, D, temperature, w);
    MVF_CHECK_LAUNCH();
  } else if (quirk) {
    launch_k(scl_cross_kernel, cross_grid, 256, smem, st, embs, D, T2, temperature, w.valid, w.counts, w.masked, w.counts + 1,
                                                    nullptr, 1.f, nullptr, 1e-6f, 0, w.zext, nullptr);
    MVF_CHECK_LAUNCH();
  }
  if (batch) {
    launch_k(scl_cross_kernel, cross_grid, 256, smem, st, embs, D, T2, temperature, w.valid, w.counts, w.valid, w.counts,
                                                    nullptr, 1.f, nullptr, 1.f, 1, w.zext, nullptr);
    MVF_CHECK_LAUNCH();
  }

  // own-pair part: one fused CTA per pair when the pair fits in shared memory, else the two row-warp passes
  const size_t fsmem = scl_fused_smem(T, D);
  static int fused_on = -1;
  if (fused_on < 0) {
    const char* e = getenv("MVF_SCL_FUSED");   // 0: row-warp kernels, 1: fused v1, 2 (default): fused v2
    fused_on = e ? atoi(e) : 2;
  }
  const size_t f2smem = scl_fused2_smem(T, D);
  if (scl_pair_tc_enabled(T, D)) {
    // prototype (MVF_SCL_TC=1): tensor-core per-pair kernel, see scl_tc.cu
    MVF_TRY(scl_pair_tc(embs, seq_lens, steps, masks, Bv, T, D, temperature, 2.f * label_variance, w.M, w.zext, w.c, loss_out,
                        d_embs, st));
  } else if (fused_on >= 2 && f2smem <= 200 * 1024) {
    if (T * T >= 2048) {
      static size_t configured = 0;
      if (f2smem > configured) {
        MVF_CHECK_CUDA(cudaFuncSetAttribute(scl_pair_fused2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f2smem));
        configured = f2smem;
      }
      launch_k(scl_pair_fused2_kernel<256>, Bv, 256, f2smem, st, embs, seq_lens, steps, masks, T, D, temperature, 2.f * label_variance,
                                                          w, loss_out, d_embs);
    } else {
      static size_t configured = 0;
      if (f2smem > configured) {
        MVF_CHECK_CUDA(cudaFuncSetAttribute(scl_pair_fused2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f2smem));
        configured = f2smem;
      }
      launch_k(scl_pair_fused2_kernel<128>, Bv, 128, f2smem, st, embs, seq_lens, steps, masks, T, D, temperature, 2.f * label_variance,
                                                          w, loss_out, d_embs);
    }
    MVF_CHECK_LAUNCH();
  } else if (fused_on && fsmem <= 200 * 1024) {
    if (T * T >= 2048) {
      static size_t configured = 0;
      if (fsmem > configured) {
        MVF_CHECK_CUDA(cudaFuncSetAttribute(scl_pair_fused_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        configured = fsmem;
      }
      launch_k(scl_pair_fused_kernel<256>, Bv, 256, fsmem, st, embs, seq_lens, steps, masks, T, D, temperature, 2.f * label_variance,
                                                         w, loss_out, d_embs);
    } else {
      static size_t configured = 0;
      if (fsmem > configured) {
        MVF_CHECK_CUDA(cudaFuncSetAttribute(scl_pair_fused_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        configured = fsmem;
      }
      launch_k(scl_pair_fused_kernel<128>, Bv, 128, fsmem, st, embs, seq_lens, steps, masks, T, D, temperature, 2.f * label_variance,
                                                         w, loss_out, d_embs);
    }
    MVF_CHECK_LAUNCH();
  } else {
    const int pair_grid = Bv * 2 * cdiv(T, 8);
    launch_k(scl_pair_kernel<0>, pair_grid, 256, smem, st, embs, seq_lens, steps, masks, T, D, temperature,
                                                     2.f * label_variance, w, loss_out, nullptr);
    MVF_CHECK_LAUNCH();
    if (d_embs) {
      launch_k(scl_pair_kernel<1>, pair_grid, 256, smem, st, embs, seq_lens, steps, masks, T, D, temperature,
                                                       2.f * label_variance, w, loss_out, d_embs);
      MVF_CHECK_LAUNCH();
    }
  }
  if (d_embs) {
    if (use_mex) {
      // both masked-column gradient terms from the stored matrix, one launch (roles by block index)
      // role A: valid output rows with the few masked tiles in one share each; role B: the masked output rows, whose many
      // valid inner tiles are what the 8 shares of blockIdx.y split (role A CTAs with blockIdx.y past their tiles exit)
      const int gridA = cx, gridB = cx < 64 ? cx : 64;
      launch_k(scl_mgrad_kernel, dim3(gridA + gridB, 8), 256, smem, st, embs, D, temperature, w, gridA, d_embs);
      MVF_CHECK_LAUNCH();
    } else if (quirk) {
      // rows valid i, columns masked k: dE_i += c_i 1e-6 e^{l_ik} e_k / tau
      launch_k(scl_cross_kernel, cross_grid, 256, smem, st, embs, D, T2, temperature, w.valid, w.counts, w.masked,
                                                      w.counts + 1, w.c, 1.f, nullptr, 1e-6f, 0, nullptr, d_embs);
      MVF_CHECK_LAUNCH();
      // rows masked k, columns valid i: dE_k += 1e-6 sum_i c_i e^{l_ik} e_i / tau
      launch_k(scl_cross_kernel, cross_grid_tall, 256, smem, st, embs, D, T2, temperature, w.masked, w.counts + 1, w.valid,
                                                           w.counts, nullptr, 1e-6f, w.c, 1.f, 0, nullptr, d_embs);
      MVF_CHECK_LAUNCH();
    }
    if (batch) {
      launch_k(scl_cross_kernel, cross_grid, 256, smem, st, embs, D, T2, temperature, w.valid, w.counts, w.valid, w.counts,
                                                      w.c, 1.f, nullptr, 1.f, 1, nullptr, d_embs);
      MVF_CHECK_LAUNCH();
      launch_k(scl_cross_kernel, cross_grid, 256, smem, st, embs, D, T2, temperature, w.valid, w.counts, w.valid, w.counts,
                                                      nullptr, 1.f, w.c, 1.f, 1, nullptr, d_embs);
      MVF_CHECK_LAUNCH();
    }
  }
  return MVF_OK;
}

}  // namespace mvf

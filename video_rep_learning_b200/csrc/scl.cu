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
};

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
  extern __shared__ __align__(16) float sm[];
  const int Dp = D + 4;
  float* tile = sm;                 // [32][Dp]
  float* er = tile + 32 * Dp;       // [8][D]
  float* ccs = er + 8 * D;          // [32]
  int* cvid = (int*)(ccs + 32);     // [32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrows = *row_cnt, ncols = *col_cnt;
  const int rslot = blockIdx.x * 8 + warp;
  if (blockIdx.x * 8 >= nrows) return;  // whole CTA idle (uniform)
  const bool ractive = rslot < nrows;
  const int r = ractive ? row_idx[rslot] : 0;
  const int rvid = r / T2;
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
  if (!ractive) return;
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

  scl_prep_kernel<<<1, 1024, 0, st>>>(masks, N, w, loss_out);
  MVF_CHECK_LAUNCH();

  const size_t smem = ((size_t)32 * (D + 4) + 8 * D + 64) * sizeof(float);
  // the column list is split over blockIdx.y: the row/column counts live on the device (valid vs masked frames), so the
  // grid is sized for the worst case and short lists simply leave CTAs idle; 8-way split keeps every pass to a few
  // tiles per CTA at training batch sizes
  int cy = 8;
  if ((int64_t)cdiv(N, 8) * cy > 65535 * 4) cy = 2;
  const dim3 cross_grid(cdiv(N, 8), cy);
  const int T2 = 2 * T;
  const bool batch = negative_type == MVF_NEG_BATCH_NOSELF;
  // Z extras
  if (quirk) {
    scl_cross_kernel<<<cross_grid, 256, smem, st>>>(embs, D, T2, temperature, w.valid, w.counts, w.masked, w.counts + 1,
                                                    nullptr, 1.f, nullptr, 1e-6f, 0, w.zext, nullptr);
    MVF_CHECK_LAUNCH();
  }
  if (batch) {
    scl_cross_kernel<<<cross_grid, 256, smem, st>>>(embs, D, T2, temperature, w.valid, w.counts, w.valid, w.counts,
                                                    nullptr, 1.f, nullptr, 1.f, 1, w.zext, nullptr);
    MVF_CHECK_LAUNCH();
  }

  const int pair_grid = Bv * 2 * cdiv(T, 8);
  scl_pair_kernel<0><<<pair_grid, 256, smem, st>>>(embs, seq_lens, steps, masks, T, D, temperature,
                                                   2.f * label_variance, w, loss_out, nullptr);
  MVF_CHECK_LAUNCH();
  if (d_embs) {
    scl_pair_kernel<1><<<pair_grid, 256, smem, st>>>(embs, seq_lens, steps, masks, T, D, temperature,
                                                     2.f * label_variance, w, loss_out, d_embs);
    MVF_CHECK_LAUNCH();
    if (quirk) {
      // rows valid i, columns masked k: dE_i += c_i 1e-6 e^{l_ik} e_k / tau
      scl_cross_kernel<<<cross_grid, 256, smem, st>>>(embs, D, T2, temperature, w.valid, w.counts, w.masked,
                                                      w.counts + 1, w.c, 1.f, nullptr, 1e-6f, 0, nullptr, d_embs);
      MVF_CHECK_LAUNCH();
      // rows masked k, columns valid i: dE_k += 1e-6 sum_i c_i e^{l_ik} e_i / tau
      scl_cross_kernel<<<cross_grid, 256, smem, st>>>(embs, D, T2, temperature, w.masked, w.counts + 1, w.valid,
                                                      w.counts, nullptr, 1e-6f, w.c, 1.f, 0, nullptr, d_embs);
      MVF_CHECK_LAUNCH();
    }
    if (batch) {
      scl_cross_kernel<<<cross_grid, 256, smem, st>>>(embs, D, T2, temperature, w.valid, w.counts, w.valid, w.counts,
                                                      w.c, 1.f, nullptr, 1.f, 1, nullptr, d_embs);
      MVF_CHECK_LAUNCH();
      scl_cross_kernel<<<cross_grid, 256, smem, st>>>(embs, D, T2, temperature, w.valid, w.counts, w.valid, w.counts,
                                                      nullptr, 1.f, w.c, 1.f, 1, nullptr, d_embs);
      MVF_CHECK_LAUNCH();
    }
  }
  return MVF_OK;
}

}  // namespace mvf

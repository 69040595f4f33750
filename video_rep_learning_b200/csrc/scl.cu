// Sequence Contrastive Loss, forward + gradient fused (algos/scl.py:52-105): workspace, index preparation and the launch
// sequence.  The arithmetic lives in scl_mma.cu (mma.sync on bf16 hi / lo operand splits).
//
// The reference builds ~10 N x N fp32 temporaries (N = Bv*2*T) with Python loops over the batch.  Here:
//   * the useful work is per video pair: one T x T logit block S = E0 E1^T / tau serves both view directions;
//   * a row i = (video, view, frame) only interacts with its partner block (own video, other view), plus
//     the "extra" columns that enter its partition sum Z_i:
//       - every MASKED frame of the local batch with weight 1e-6 (weight.masked_fill_(mask==0, 1e-6) runs
//         after the single/noself zeroing, scl.py:80)                                   [quirk = 1]
//       - for NEGATIVE_TYPE batch_noself every valid frame of the OTHER videos with weight 1 (scl.py:74-79).
// Launch sequence (one stream, no host sync; M = sum(mask) and the list sizes stay on the device):
//   prep  ->  [cross sums: Z extras]  ->  pair kernel (loss, c, own-pair dE)  ->  [cross gradients: dE extras]
// with the bracketed launches only when quirk / batch_noself ask for them (they return at once when a list is empty).
// Nothing of size T x T or N x N ever reaches HBM.
//
// Closed form, direction with rows i and partner columns j, mm_ij = m_i m_j:
//   y_ij = pw_ij / sum_j pw_ij,  pw_ij = mm_ij exp(-d_ij^2 / (2 var)),  d_ij = |fl(fl(s_i / L_i) * L_j) - s_j|
//   Z_i = sum_j mm_ij e^{l_ij} + zext_i,  p = e^{l}/Z,  q = p + 1e-6,  loss += mm y (log y - log q) / M
//   r = p/q,  g_i = sum_j mm y r,  dloss/dl_ij = mm (p g_i - y r) / M,  extras: w c_i e^{l_ik},  c_i = g_i/(Z_i M)
#include "kernels.cuh"
#include "scl_ws.cuh"

namespace mvf {

size_t scl_ws_layout(int N, SclWs* w, char* base) {
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
  if (w) { w->M = M; w->Z = Z; w->g = g; w->den = den; w->c = c; w->zext = zext; w->counts = counts; w->valid = valid; w->masked = masked; w->chunk = chunk; }
  return off;
}
size_t scl_ws_bytes(int Bv, int T, int D) {
  (void)D;
  return scl_ws_layout(Bv * 2 * T, nullptr, nullptr);
}

// ---- prep: M = sum(mask), ordered index lists of valid / masked rows, zeroed accumulators --------------------
// N <= 8192: one CTA, each of its 32 warps owns a contiguous slice of rows; ballot counts, one prefix over the 32 warp
// totals, then every warp writes its slice of the two ordered lists (deterministic, no host round trip).
constexpr int SCL_PREP_SMALL = 8192;
__global__ void __launch_bounds__(1024) scl_prep_kernel(const float* __restrict__ masks, int N, SclWs w, float* loss_out) {
  pdl_entry();
  __shared__ int cnt[32];
  __shared__ float sums[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = ((N + 31) / 32 + 31) / 32 * 32;            // rows per warp, a multiple of 32
  const int r0 = min(N, warp * per), r1 = min(N, r0 + per);
  int nval = 0;
  float s = 0.f;
  for (int i0 = r0; i0 < r1; i0 += 32) {
    const int i = i0 + lane;
    const float m = i < r1 ? masks[i] : 0.f;
    if (i < r1) { w.zext[i] = 0.f; w.c[i] = 0.f; }
    nval += __popc(__ballot_sync(0xffffffffu, m != 0.f));
    s += m;
  }
  s = warp_sum(s);
  if (lane == 0) { cnt[warp] = nval; sums[warp] = s; }
  __syncthreads();
  const int before = __reduce_add_sync(0xffffffffu, lane < warp ? cnt[lane] : 0);
  int bv = before, bm = r0 - before;
  for (int i0 = r0; i0 < r1; i0 += 32) {
    const int i = i0 + lane;
    const bool in = i < r1;
    const bool val = in && masks[i] != 0.f;
    const unsigned bal = __ballot_sync(0xffffffffu, val), inb = __ballot_sync(0xffffffffu, in);
    const unsigned lt = (1u << lane) - 1u;
    if (val) w.valid[bv + __popc(bal & lt)] = i;
    else if (in) w.masked[bm + __popc(~bal & inb & lt)] = i;
    bv += __popc(bal);
    bm += __popc(~bal & inb);
  }
  if (warp == 0) {
    const int total = __reduce_add_sync(0xffffffffu, cnt[lane]);
    const float msum = warp_sum(sums[lane]);
    if (lane == 0) {
      *w.M = msum;
      *loss_out = 0.f;
      w.counts[0] = total;
      w.counts[1] = N - total;
    }
  }
}

// ---- prep for large batches: the same outputs from three launches ---------------------------------------------------
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

int scl_fwd_bwd(const float* embs, const int64_t* seq_lens, const int64_t* steps, const float* masks, int Bv, int T,
                int D, float temperature, float label_variance, int negative_type, int quirk, float* loss_out,
                float* d_embs, void* ws, size_t ws_bytes, cudaStream_t st) {
  MVF_REQUIRE(embs && seq_lens && steps && masks && loss_out && ws, MVF_ERR_BAD_ARG, "scl: null pointer");
  MVF_REQUIRE(Bv > 0 && T > 0 && D > 0, MVF_ERR_BAD_ARG, "scl: bad shape Bv=%d T=%d D=%d", Bv, T, D);
  MVF_REQUIRE(scl_mma_supported(T, D), MVF_ERR_UNSUPPORTED,
              "scl: %d frames x %d channels (at most 256 frames per view, 256 channels, channels a multiple of 4)", T, D);
  MVF_REQUIRE((((uintptr_t)embs) & 15) == 0, MVF_ERR_ALIGN, "scl: embeddings must be 16-byte aligned");
  MVF_REQUIRE(negative_type == MVF_NEG_SINGLE_NOSELF || negative_type == MVF_NEG_BATCH_NOSELF, MVF_ERR_UNSUPPORTED,
              "scl: negative_type %d", negative_type);
  const int64_t N64 = (int64_t)Bv * 2 * T;
  MVF_REQUIRE(N64 < (1ll << 30), MVF_ERR_BAD_ARG, "scl: batch too large");
  const int N = (int)N64;
  SclWs w;
  const size_t need = scl_ws_layout(N, &w, (char*)ws);
  MVF_REQUIRE(ws_bytes >= need, MVF_ERR_WORKSPACE, "scl: workspace %zu < %zu bytes", ws_bytes, need);

  if (N <= SCL_PREP_SMALL) {
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

  const bool batch = negative_type == MVF_NEG_BATCH_NOSELF;
  const int T2 = 2 * T;
  // Z extras: masked frames of the batch (weight 1e-6) and, for batch_noself, the valid frames of the other videos
  SclCrossJobs sums;
  sums.n = 0;
  if (quirk) sums.job[sums.n++] = SclCrossJob{w.valid, w.counts, w.masked, w.counts + 1, nullptr, nullptr, 1e-6f, 0};
  if (batch) sums.job[sums.n++] = SclCrossJob{w.valid, w.counts, w.valid, w.counts, nullptr, nullptr, 1.f, 1};
  if (sums.n) MVF_TRY(scl_cross_mma(embs, N, T2, D, temperature, sums, 0, w.zext, nullptr, st));

  MVF_TRY(scl_pair_mma(embs, seq_lens, steps, masks, Bv, T, D, temperature, 2.f * label_variance, w, sums.n > 0, loss_out,
                       d_embs, st));

  if (d_embs && sums.n) {
    SclCrossJobs gr;
    gr.n = 0;
    if (quirk) {
      // rows valid i, columns masked k: dE_i += c_i 1e-6 e^{l_ik} e_k / tau;  rows masked k, columns valid i: dE_k += ...
      gr.job[gr.n++] = SclCrossJob{w.valid, w.counts, w.masked, w.counts + 1, w.c, nullptr, 1e-6f, 0};
      gr.job[gr.n++] = SclCrossJob{w.masked, w.counts + 1, w.valid, w.counts, nullptr, w.c, 1e-6f, 0};
    }
    if (batch) {
      gr.job[gr.n++] = SclCrossJob{w.valid, w.counts, w.valid, w.counts, w.c, nullptr, 1.f, 1};
      gr.job[gr.n++] = SclCrossJob{w.valid, w.counts, w.valid, w.counts, nullptr, w.c, 1.f, 1};
    }
    MVF_TRY(scl_cross_mma(embs, N, T2, D, temperature, gr, 1, nullptr, d_embs, st));
  }
  return MVF_OK;
}

}  // namespace mvf

// Memory-bound glue kernels of the head: parameter packing, positional encoding, LayerNorm (+residual
// +dropout), BatchNorm1d (+ReLU +dropout) with split statistics (so the statistics can be all-reduced across
// ranks between two launches), column sums for bias gradients, entity reduction, L2 normalisation.
// All fp32 math; outputs optionally bf16 when they feed a tensor-core GEMM.
#include <stdlib.h>

#include "kernels.cuh"

namespace mvf {

// =========================================== pack / unpack ===========================================
constexpr int PACK_CHUNK = 64;   // entries per launch (kernel parameter table: 2.6 KB)
struct PackTable {
  int n;
  PackEntry e[PACK_CHUNK];
};
struct UnpackTable {
  int n;
  UnpackEntry e[PACK_CHUNK];
};

__global__ void pack_kernel(const PackTable tab) {
  pdl_entry();
  const PackEntry& en = tab.e[blockIdx.y];
  const int64_t total = (int64_t)en.rows * en.ld_dst;
  if (en.perm_n <= 1 && en.dst_bf16 == 0 && en.ld_dst == en.cols && (total & 3) == 0 && ((((uintptr_t)en.src) | ((uintptr_t)en.dst)) & 15) == 0) {
    // plain fp32 copy of a dense matrix: 16-byte vectors, no index arithmetic
    const float4* s4 = reinterpret_cast<const float4*>(en.src);
    float4* d4 = reinterpret_cast<float4*>(en.dst);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (total >> 2); i += (int64_t)gridDim.x * blockDim.x) d4[i] = s4[i];
    return;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / en.ld_dst), c = (int)(i - (int64_t)r * en.ld_dst);
    int rs = r, cs = c;
    if (en.perm_n > 1) {   // interleaved source rows (FWBPooling: output channel c*E + e feeds entity e, channel c)
      if (en.rows > 1) rs = (r % en.perm_n) * (en.rows / en.perm_n) + r / en.perm_n;
      else if (c < en.cols) cs = (c % en.perm_n) * (en.cols / en.perm_n) + c / en.perm_n;
    }
    float v = c < en.cols ? en.src[(int64_t)rs * en.cols + cs] : 0.f;
    if (en.dst_bf16 == 2) {
      // "pre-split" container for the bf16x3 GEMM: every 32-float block of a row holds [hi(32) | lo(32)] bf16
      const bf16 hi = __float2bfloat16_rn(v);
      const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      bf16* row = (bf16*)en.dst + (int64_t)r * en.ld_dst * 2;
      row[(c >> 5) * 64 + (c & 31)] = hi;
      row[(c >> 5) * 64 + 32 + (c & 31)] = lo;
    } else if (en.dst_bf16) ((bf16*)en.dst)[i] = __float2bfloat16_rn(v);
    else ((float*)en.dst)[i] = v;
  }
}

int pack_params(const PackEntry* entries, int n, cudaStream_t st) {
  for (int base = 0; base < n; base += PACK_CHUNK) {
    PackTable tab;
    tab.n = (n - base < PACK_CHUNK) ? n - base : PACK_CHUNK;
    for (int i = 0; i < tab.n; ++i) tab.e[i] = entries[base + i];
    launch_k(pack_kernel, dim3(148, tab.n), 256, 0, st, tab);
    MVF_CHECK_LAUNCH();
  }
  return MVF_OK;
}

__global__ void unpack_kernel(const UnpackTable tab, float scale) {
  pdl_entry();
  const UnpackEntry& en = tab.e[blockIdx.y];
  const int64_t total = (int64_t)en.rows * en.cols;
  if (en.perm_n <= 1 && en.ld_src == en.cols && (total & 3) == 0 && ((((uintptr_t)en.src) | ((uintptr_t)en.dst)) & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(en.src);
    float4* d4 = reinterpret_cast<float4*>(en.dst);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (total >> 2); i += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = s4[i];
      d4[i] = make_float4(scale * v.x, scale * v.y, scale * v.z, scale * v.w);
    }
    return;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / en.cols), c = (int)(i - (int64_t)r * en.cols);
    if (en.perm_n > 1) {
      if (en.rows > 1) r = (r % en.perm_n) * (en.rows / en.perm_n) + r / en.perm_n;
      else c = (c % en.perm_n) * (en.cols / en.perm_n) + c / en.perm_n;
    }
    en.dst[i] = scale * en.src[(int64_t)r * en.ld_src + c];
  }
}

int unpack_grads(const UnpackEntry* entries, int n, float scale, cudaStream_t st) {
  for (int base = 0; base < n; base += PACK_CHUNK) {
    UnpackTable tab;
    tab.n = (n - base < PACK_CHUNK) ? n - base : PACK_CHUNK;
    for (int i = 0; i < tab.n; ++i) tab.e[i] = entries[base + i];
    launch_k(unpack_kernel, dim3(148, tab.n), 256, 0, st, tab, scale);
    MVF_CHECK_LAUNCH();
  }
  return MVF_OK;
}

// ======================================== positional encoding ========================================
// models/utils.py:113-126: sin on even channels, cos on odd, exponent uses the channel index itself;
// positions 0..T-1 or numpy.linspace(0, train-1, T) when T != train length.  float64 like numpy.
__global__ void posenc_table_kernel(float* table, int T, int H, int train_frames) {
  pdl_entry();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * H) return;
  int t = i / H, c = i % H;
  double pos;
  if (T == train_frames) pos = (double)t;
  else if (T == 1) pos = 0.0;
  else {
    double step = (double)(train_frames - 1) / (double)(T - 1);
    pos = (t == T - 1) ? (double)(train_frames - 1) : (double)t * step;
  }
  double ang = pos / pow(10000.0, (double)c / (double)H);
  table[i] = (float)((c & 1) ? cos(ang) : sin(ang));
}
int posenc_table(float* table, int T, int H, int train_frames, cudaStream_t st) {
  launch_k(posenc_table_kernel, cdiv((int64_t)T * H, 256), 256, 0, st, table, T, H, train_frames);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

__global__ void posenc_add_kernel(const float* __restrict__ h3, const float* __restrict__ table, float* __restrict__ z,
                                  int BV, int T, int E, int H, float p, float inv_keep, DropSeed seed) {
  pdl_entry();
  int64_t total = (int64_t)BV * E * T * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % H);
    int64_t row = i / H;  // b*E*T + e*T + t
    int t = (int)(row % T);
    int e = (int)((row / T) % E);
    int b = (int)(row / ((int64_t)T * E));
    float v = h3[(((int64_t)b * T + t) * E + e) * H + c] + table[t * H + c];
    if (p > 0.f) v *= drop_scale(seed, SITE_POS, (uint64_t)i, p, inv_keep);
    z[i] = v;
  }
}
// 16-byte version (H % 4 == 0, < 2^31 elements): no 64-bit divisions, four elements per thread
__device__ __forceinline__ uint64_t seed_value(DropSeed s) { return s.base + (s.dev ? __ldg(s.dev) : 0ull); }
__device__ __forceinline__ float4 drop4(float4 v, uint64_t sd, int site, uint64_t idx0, float p, float inv_keep) {
  v.x *= drop_scale(sd, site, idx0, p, inv_keep);
  v.y *= drop_scale(sd, site, idx0 + 1, p, inv_keep);
  v.z *= drop_scale(sd, site, idx0 + 2, p, inv_keep);
  v.w *= drop_scale(sd, site, idx0 + 3, p, inv_keep);
  return v;
}
__global__ void __launch_bounds__(256)
posenc_add_v4_kernel(const float4* __restrict__ h3, const float4* __restrict__ table, float4* __restrict__ z, int BV, int T,
                     int E, int H4, float p, float inv_keep, DropSeed seed) {
  pdl_entry();
  const int total4 = BV * E * T * H4;
  const uint64_t sd = seed_value(seed);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += gridDim.x * blockDim.x) {
    const int row = q / H4, c4 = q - row * H4;   // row = b*E*T + e*T + t
    const int be = row / T, t = row - be * T;
    const int b = be / E, e = be - b * E;
    const float4 a = h3[((b * T + t) * E + e) * H4 + c4], tb = table[t * H4 + c4];
    float4 v = make_float4(a.x + tb.x, a.y + tb.y, a.z + tb.z, a.w + tb.w);
    if (p > 0.f) v = drop4(v, sd, SITE_POS, (uint64_t)q * 4, p, inv_keep);
    z[q] = v;
  }
}
__global__ void __launch_bounds__(256)
posenc_bwd_v4_kernel(const float4* __restrict__ dz, float4* __restrict__ dh3, int BV, int T, int E, int H4, float p,
                     float inv_keep, DropSeed seed) {
  pdl_entry();
  const int total4 = BV * E * T * H4;
  const uint64_t sd = seed_value(seed);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += gridDim.x * blockDim.x) {
    const int row = q / H4, c4 = q - row * H4;
    const int be = row / T, t = row - be * T;
    const int b = be / E, e = be - b * E;
    float4 v = dz[q];
    if (p > 0.f) v = drop4(v, sd, SITE_POS, (uint64_t)q * 4, p, inv_keep);
    dh3[((b * T + t) * E + e) * H4 + c4] = v;
  }
}
static bool vec4_ok(int64_t total, int64_t inner, const void* a, const void* b, const void* c = nullptr) {
  return (inner & 3) == 0 && total < (1ll << 31) && ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c)) & 15) == 0;
}
static int grid_for(int64_t work, int cap = 2368) { return (int)((work + 255) / 256 < cap ? (work + 255) / 256 : cap); }

int posenc_add(const float* h3, const float* table, float* z, int BV, int T, int E, int H, float p, DropSeed seed,
               cudaStream_t st) {
  int64_t total = (int64_t)BV * E * T * H;
  if (vec4_ok(total, H, h3, table, z)) {
    launch_k(posenc_add_v4_kernel, grid_for(total / 4), 256, 0, st, (const float4*)h3, (const float4*)table, (float4*)z, BV, T, E,
             H / 4, p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  }
  int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  launch_k(posenc_add_kernel, grid, 256, 0, st, h3, table, z, BV, T, E, H, p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

template <typename TO>
__global__ void posenc_bwd_kernel(const float* __restrict__ dz, TO* __restrict__ dh3, int BV, int T, int E, int H,
                                  float p, float inv_keep, DropSeed seed) {
  pdl_entry();
  int64_t total = (int64_t)BV * E * T * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % H);
    int64_t row = i / H;
    int t = (int)(row % T);
    int e = (int)((row / T) % E);
    int b = (int)(row / ((int64_t)T * E));
    float v = dz[i];
    if (p > 0.f) v *= drop_scale(seed, SITE_POS, (uint64_t)i, p, inv_keep);
    dh3[(((int64_t)b * T + t) * E + e) * H + c] = from_f<TO>(v);
  }
}
int posenc_bwd(int dtype_out, const float* dz, void* dh3, int BV, int T, int E, int H, float p, DropSeed seed,
               cudaStream_t st) {
  int64_t total = (int64_t)BV * E * T * H;
  int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  float ik = p > 0.f ? 1.f / (1.f - p) : 1.f;
  if (dtype_out == MVF_F32 && vec4_ok(total, H, dz, dh3)) {
    launch_k(posenc_bwd_v4_kernel, grid_for(total / 4), 256, 0, st, (const float4*)dz, (float4*)dh3, BV, T, E, H / 4, p, ik, seed);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  }
  if (dtype_out == MVF_BF16) launch_k(posenc_bwd_kernel<bf16>, grid, 256, 0, st, dz, (bf16*)dh3, BV, T, E, H, p, ik, seed);
  else launch_k(posenc_bwd_kernel<float>, grid, 256, 0, st, dz, (float*)dh3, BV, T, E, H, p, ik, seed);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// ============================================ LayerNorm =============================================
// One warp per row.  z_out = z_in + drop(o); two-pass mean/variance (biased, eps inside the sqrt).
template <typename TO>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ z_in, const float* __restrict__ o,
                                                     float* __restrict__ z_out, TO* __restrict__ r,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     int64_t rows, int H, float eps, float p, float inv_keep,
                                                     DropSeed seed, int site) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* zi = z_in + row * H;
  float* zo = z_out + row * H;
  float s = 0.f;
  for (int c = lane; c < H; c += 32) {
    float v = zi[c];
    if (o) {
      float ov = o[row * H + c];
      if (p > 0.f) ov *= drop_scale(seed, site, (uint64_t)(row * H + c), p, inv_keep);
      v += ov;
    }
    if (o || zo != zi) zo[c] = v;
    s += v;
  }
  if (r == nullptr) return;
  s = warp_sum(s);
  const float mean = s / (float)H;
  float ss = 0.f;
  for (int c = lane; c < H; c += 32) {
    float dv = zo[c] - mean;
    ss += dv * dv;
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / (float)H + eps);
  for (int c = lane; c < H; c += 32) r[row * H + c] = from_f<TO>((zo[c] - mean) * rstd * gamma[c] + beta[c]);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// Same, for H = 128 * V4 with 16-byte aligned rows: the row lives in registers (V4 float4 per lane), every array is touched
// once with 16-byte accesses (the generic kernel re-reads z_out from memory for the variance and for the output).
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a);
  u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
template <typename TO, int V4>
__global__ void __launch_bounds__(256) ln_fwd_vec_kernel(const float* __restrict__ z_in, const float* __restrict__ o,
                                                         float* __restrict__ z_out, TO* __restrict__ r,
                                                         float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         int64_t rows, float eps, float p, float inv_keep, DropSeed seed,
                                                         int site) {
  pdl_entry();
  constexpr int H = 128 * V4;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* zi = z_in + row * H;
  float* zo = z_out + row * H;
  float4 v[V4];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < V4; ++k) {
    const int c = 4 * (lane + 32 * k);
    v[k] = *reinterpret_cast<const float4*>(zi + c);
    if (o) {
      float4 ov = *reinterpret_cast<const float4*>(o + row * H + c);
      if (p > 0.f) {
        const uint64_t idx = (uint64_t)(row * H + c);
        ov.x *= drop_scale(seed, site, idx, p, inv_keep);
        ov.y *= drop_scale(seed, site, idx + 1, p, inv_keep);
        ov.z *= drop_scale(seed, site, idx + 2, p, inv_keep);
        ov.w *= drop_scale(seed, site, idx + 3, p, inv_keep);
      }
      v[k].x += ov.x; v[k].y += ov.y; v[k].z += ov.z; v[k].w += ov.w;
    }
    if (o || zo != zi) *reinterpret_cast<float4*>(zo + c) = v[k];
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  if (r == nullptr) return;
  s = warp_sum(s);
  const float mean = s / (float)H;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < V4; ++k) {
    const float a = v[k].x - mean, b = v[k].y - mean, c2 = v[k].z - mean, d = v[k].w - mean;
    ss += (a * a + b * b) + (c2 * c2 + d * d);
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / (float)H + eps);
#pragma unroll
  for (int k = 0; k < V4; ++k) {
    const int c = 4 * (lane + 32 * k);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    store4(r + row * H + c, make_float4((v[k].x - mean) * rstd * g.x + b.x, (v[k].y - mean) * rstd * g.y + b.y,
                                        (v[k].z - mean) * rstd * g.z + b.z, (v[k].w - mean) * rstd * g.w + b.w));
  }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// z_out = z_in + drop(o), 16 bytes per thread (the residual add behind the last encoder sub-layer: no LayerNorm follows)
__global__ void __launch_bounds__(256)
resid_add_v4_kernel(const float4* __restrict__ z_in, const float4* __restrict__ o, float4* __restrict__ z_out, int total4, float p,
                    float inv_keep, DropSeed seed, int site) {
  pdl_entry();
  const uint64_t sd = seed_value(seed);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += gridDim.x * blockDim.x) {
    float4 ov = o[q];
    if (p > 0.f) ov = drop4(ov, sd, site, (uint64_t)q * 4, p, inv_keep);
    const float4 zi = z_in[q];
    z_out[q] = make_float4(zi.x + ov.x, zi.y + ov.y, zi.z + ov.z, zi.w + ov.w);
  }
}

template <typename TO>
static bool ln_fwd_vec(const float* z_in, const float* o, float* z_out, TO* r, float* mean, float* rstd, const float* gamma,
                       const float* beta, int64_t rows, int H, float eps, float p, float ik, DropSeed seed, int site,
                       cudaStream_t st) {
  const uintptr_t al = (uintptr_t)z_in | (uintptr_t)o | (uintptr_t)z_out | (uintptr_t)r | (uintptr_t)gamma | (uintptr_t)beta;
  if (H % 128 != 0 || H > 1024 || (al & 15) != 0) return false;
  const int grid = cdiv(rows, 8);
#define MVF_LNF(V) launch_k(ln_fwd_vec_kernel<TO, V>, grid, 256, 0, st, z_in, o, z_out, r, mean, rstd, gamma, beta, rows, eps, p, ik, seed, site)
  switch (H / 128) {
    case 1: MVF_LNF(1); return true;
    case 2: MVF_LNF(2); return true;
    case 3: MVF_LNF(3); return true;
    case 4: MVF_LNF(4); return true;
    case 6: MVF_LNF(6); return true;
    case 8: MVF_LNF(8); return true;
    default: return false;
  }
#undef MVF_LNF
}

int ln_fwd(int dtype_out, const float* z_in, const float* o, float* z_out, void* r, float* mean, float* rstd,
           const float* gamma, const float* beta, int64_t rows, int H, float eps, float p, DropSeed seed, int site,
           cudaStream_t st) {
  float ik = p > 0.f ? 1.f / (1.f - p) : 1.f;
  if (r != nullptr && rows > 0) {
    const bool done = dtype_out == MVF_BF16 ? ln_fwd_vec<bf16>(z_in, o, z_out, (bf16*)r, mean, rstd, gamma, beta, rows, H, eps, p, ik, seed, site, st)
                                            : ln_fwd_vec<float>(z_in, o, z_out, (float*)r, mean, rstd, gamma, beta, rows, H, eps, p, ik, seed, site, st);
    if (done) {
      MVF_CHECK_LAUNCH();
      return MVF_OK;
    }
  }
  if (r == nullptr && o != nullptr && rows > 0 && vec4_ok(rows * H, H, z_in, o, z_out)) {   // residual add only
    launch_k(resid_add_v4_kernel, grid_for(rows * H / 4), 256, 0, st, (const float4*)z_in, (const float4*)o, (float4*)z_out,
             (int)(rows * H / 4), p, ik, seed, site);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  }
  int grid = cdiv(rows, 8);
  if (dtype_out == MVF_BF16)
    launch_k(ln_fwd_kernel<bf16>, grid, 256, 0, st, z_in, o, z_out, (bf16*)r, mean, rstd, gamma, beta, rows, H, eps, p, ik,
                                              seed, site);
  else
    launch_k(ln_fwd_kernel<float>, grid, 256, 0, st, z_in, o, z_out, (float*)r, mean, rstd, gamma, beta, rows, H, eps, p, ik,
                                               seed, site);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// dz_out = dz_in + rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dr*gamma; dgamma += dr*xhat; dbeta += dr
// One warp per row, each lane owns VPL = H/32 fixed columns: the row is read once into registers, the per-column
// gamma/beta gradients accumulate in registers across the warp's rows, then warps are merged through shared memory
// and one atomic per column per CTA reaches the gradient buffer.
template <int VPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dr, const float* __restrict__ z,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ dz_in,
                                                     float* __restrict__ dz_out, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, int64_t rows, int H,
                                                     float* __restrict__ drop_out, float p, float inv_keep, DropSeed seed,
                                                     int site) {
  pdl_entry();
  extern __shared__ float acc[];  // [8 warps][2][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gam[VPL], ag[VPL], ab[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    gam[k] = gamma[lane + 32 * k];
    ag[k] = 0.f;
    ab[k] = 0.f;
  }
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    const float m = mean[row], rs = rstd[row];
    float d[VPL], xh[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = lane + 32 * k;
      d[k] = dr[row * H + c];
      xh[k] = (z[row * H + c] - m) * rs;
      const float g = d[k] * gam[k];
      s1 += g;
      s2 += g * xh[k];
      ag[k] += d[k] * xh[k];
      ab[k] += d[k];
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int c = lane + 32 * k;
      float v = rs * (d[k] * gam[k] - s1 - xh[k] * s2);
      if (dz_in) v += dz_in[row * H + c];
      dz_out[row * H + c] = v;
      // the gradient that enters the next residual branch is dz masked by that branch's dropout: written here instead of
      // by a separate dropout_cast launch
      if (drop_out) drop_out[row * H + c] = p > 0.f ? v * drop_scale(seed, site, (uint64_t)(row * H + c), p, inv_keep) : v;
    }
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    acc[(size_t)warp * 2 * H + lane + 32 * k] = ag[k];
    acc[(size_t)warp * 2 * H + H + lane + 32 * k] = ab[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float g = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      g += acc[(size_t)w * 2 * H + c];
      b += acc[(size_t)w * 2 * H + H + c];
    }
    atomicAdd(dgamma + c, g);
    atomicAdd(dbeta + c, b);
  }
}

// Same for H = 128 * V4 with 16-byte aligned rows: lane owns 4 consecutive columns per float4 (16-byte accesses).
template <int V4>
__global__ void __launch_bounds__(256) ln_bwd_vec_kernel(const float* __restrict__ dr, const float* __restrict__ z,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ dz_in,
                                                         float* __restrict__ dz_out, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, int64_t rows,
                                                         float* __restrict__ drop_out, float p, float inv_keep, DropSeed seed,
                                                         int site) {
  pdl_entry();
  constexpr int H = 128 * V4;
  extern __shared__ __align__(16) float acc[];  // [8 warps][2][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 gam[V4], ag[V4], ab[V4];
#pragma unroll
  for (int k = 0; k < V4; ++k) {
    gam[k] = *reinterpret_cast<const float4*>(gamma + 4 * (lane + 32 * k));
    ag[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    const float m = mean[row], rs = rstd[row];
    float4 d[V4], xh[V4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < V4; ++k) {
      const int c = 4 * (lane + 32 * k);
      d[k] = *reinterpret_cast<const float4*>(dr + row * H + c);
      const float4 zv = *reinterpret_cast<const float4*>(z + row * H + c);
      xh[k] = make_float4((zv.x - m) * rs, (zv.y - m) * rs, (zv.z - m) * rs, (zv.w - m) * rs);
      const float gx = d[k].x * gam[k].x, gy = d[k].y * gam[k].y, gz = d[k].z * gam[k].z, gw = d[k].w * gam[k].w;
      s1 += (gx + gy) + (gz + gw);
      s2 += (gx * xh[k].x + gy * xh[k].y) + (gz * xh[k].z + gw * xh[k].w);
      ag[k].x += d[k].x * xh[k].x; ag[k].y += d[k].y * xh[k].y; ag[k].z += d[k].z * xh[k].z; ag[k].w += d[k].w * xh[k].w;
      ab[k].x += d[k].x; ab[k].y += d[k].y; ab[k].z += d[k].z; ab[k].w += d[k].w;
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
#pragma unroll
    for (int k = 0; k < V4; ++k) {
      const int c = 4 * (lane + 32 * k);
      float4 v = make_float4(rs * (d[k].x * gam[k].x - s1 - xh[k].x * s2), rs * (d[k].y * gam[k].y - s1 - xh[k].y * s2),
                             rs * (d[k].z * gam[k].z - s1 - xh[k].z * s2), rs * (d[k].w * gam[k].w - s1 - xh[k].w * s2));
      if (dz_in) {
        const float4 a = *reinterpret_cast<const float4*>(dz_in + row * H + c);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
      *reinterpret_cast<float4*>(dz_out + row * H + c) = v;
      if (drop_out) {
        float4 o = v;
        if (p > 0.f) {
          const uint64_t idx = (uint64_t)(row * H + c);
          o.x *= drop_scale(seed, site, idx, p, inv_keep);
          o.y *= drop_scale(seed, site, idx + 1, p, inv_keep);
          o.z *= drop_scale(seed, site, idx + 2, p, inv_keep);
          o.w *= drop_scale(seed, site, idx + 3, p, inv_keep);
        }
        *reinterpret_cast<float4*>(drop_out + row * H + c) = o;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < V4; ++k) {
    const int c = 4 * (lane + 32 * k);
    *reinterpret_cast<float4*>(acc + (size_t)warp * 2 * H + c) = ag[k];
    *reinterpret_cast<float4*>(acc + (size_t)warp * 2 * H + H + c) = ab[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float g = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      g += acc[(size_t)w * 2 * H + c];
      b += acc[(size_t)w * 2 * H + H + c];
    }
    atomicAdd(dgamma + c, g);
    atomicAdd(dbeta + c, b);
  }
}

int ln_bwd(const float* dr, const float* z, const float* mean, const float* rstd, const float* gamma,
           const float* dz_in, float* dz_out, float* dgamma, float* dbeta, int64_t rows, int H, cudaStream_t st,
           float* drop_out, float p, DropSeed seed, int site) {
  const float ik = p > 0.f ? 1.f / (1.f - p) : 1.f;
  int grid = cdiv(rows, 8 * 2);
  if (grid > 592) grid = 592;
  if (grid < 1) grid = 1;
  size_t smem = (size_t)8 * 2 * H * sizeof(float);
  MVF_REQUIRE(H % 32 == 0 && H <= 1024 && smem <= 48 * 1024, MVF_ERR_UNSUPPORTED,
              "ln_bwd: hidden size %d must be a multiple of 32 and <= 768", H);
  {
    const uintptr_t al = (uintptr_t)dr | (uintptr_t)z | (uintptr_t)gamma | (uintptr_t)dz_in | (uintptr_t)dz_out | (uintptr_t)drop_out;
    static int vec_on = -1;   // MVF_LN_VEC=0: scalar kernels only (A/B)
    if (vec_on < 0) {
      const char* e = getenv("MVF_LN_VEC");
      vec_on = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (vec_on && H % 128 == 0 && (al & 15) == 0) {
#define MVF_LNBV(V) launch_k(ln_bwd_vec_kernel<V>, grid, 256, smem, st, dr, z, mean, rstd, gamma, dz_in, dz_out, dgamma, dbeta, rows, drop_out, p, ik, seed, site)
      bool done = true;
      switch (H / 128) {
        case 1: MVF_LNBV(1); break;
        case 2: MVF_LNBV(2); break;
        case 3: MVF_LNBV(3); break;
        case 4: MVF_LNBV(4); break;
        case 6: MVF_LNBV(6); break;
        default: done = false; break;
      }
#undef MVF_LNBV
      if (done) {
        MVF_CHECK_LAUNCH();
        return MVF_OK;
      }
    }
  }
#define MVF_LNB(V) launch_k(ln_bwd_kernel<V>, grid, 256, smem, st, dr, z, mean, rstd, gamma, dz_in, dz_out, dgamma, dbeta, rows, H, drop_out, p, ik, seed, site)
  switch (H / 32) {
    case 1: MVF_LNB(1); break;
    case 2: MVF_LNB(2); break;
    case 3: MVF_LNB(3); break;
    case 4: MVF_LNB(4); break;
    case 6: MVF_LNB(6); break;
    case 8: MVF_LNB(8); break;
    case 12: MVF_LNB(12); break;
    case 16: MVF_LNB(16); break;
    case 24: MVF_LNB(24); break;
    default:
      set_error("ln_bwd: hidden size %d not instantiated (32,64,96,128,192,256,384,512,768)", H);
      return MVF_ERR_UNSUPPORTED;
  }
#undef MVF_LNB
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// ============================================ BatchNorm ==============================================
// Column statistics: block = 32 columns x 8 row lanes; double partials, double atomics.
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, int64_t R, int C,
                                                       double* __restrict__ sums) {
  pdl_entry();
  __shared__ double sh[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double s = 0.0, q = 0.0;
  if (c < C) {
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < R; r += (int64_t)gridDim.y * 8) {
      double v = (double)x[r * C + c];
      s += v;
      q += v * v;
    }
  }
  sh[0][ty][tx] = s;
  sh[1][ty][tx] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int j = 1; j < 8; ++j) { s += sh[0][j][tx]; q += sh[1][j][tx]; }
    atomicAdd(sums + c, s);
    atomicAdd(sums + C + c, q);
  }
}
int bn_stats(const float* x, int64_t R, int C, double* sums, cudaStream_t st) {
  int gy = cdiv(R, 8 * 16);
  if (gy > 64) gy = 64;
  if (gy < 1) gy = 1;
  launch_k(bn_stats_kernel, dim3(cdiv(C, 32), gy), 256, 0, st, x, R, C, sums);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// mean / invstd (fp32) from the (possibly all-reduced) double sums; running statistics as nn.BatchNorm1d:
// running = (1-m)*running + m*batch, with the UNBIASED batch variance.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, double n, float eps, int training,
                                   float momentum, float* __restrict__ rmean, float* __restrict__ rvar,
                                   int64_t* __restrict__ tracked, float* __restrict__ mi) {
  pdl_entry();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && tracked) *tracked += 1;
  if (c >= C) return;
  if (training) {
    double mean = sums[c] / n;
    double var = sums[C + c] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    mi[c] = (float)mean;
    mi[C + c] = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean) {
      double unb = n > 1.0 ? var * n / (n - 1.0) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
    }
  } else {
    mi[c] = rmean[c];
    mi[C + c] = 1.0f / sqrtf(rvar[c] + eps);
  }
}
int bn_finalize(const double* sums, int C, double n_global, float eps, int training, float momentum, float* rmean,
                float* rvar, int64_t* tracked, float* mi, cudaStream_t st) {
  launch_k(bn_finalize_kernel, cdiv(C, 128), 128, 0, st, sums, C, n_global, eps, training, momentum, rmean, rvar, tracked, mi);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

template <typename TO>
__global__ void bn_apply_kernel(const float* __restrict__ x, int64_t R, int C, const float* __restrict__ mi,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                TO* __restrict__ out, int64_t ld_out, float p, float inv_keep, DropSeed seed, int site) {
  pdl_entry();
  int64_t total = R * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t r = i / C;
    float y = (x[i] - mi[c]) * mi[C + c] * gamma[c] + beta[c];
    if (relu) y = fmaxf(y, 0.f);
    if (p > 0.f) y *= drop_scale(seed, site, (uint64_t)i, p, inv_keep);
    out[r * ld_out + c] = from_f<TO>(y);
  }
}
__global__ void __launch_bounds__(256)
bn_apply_v4_kernel(const float4* __restrict__ x, int R, int C4, const float* __restrict__ mi, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int relu, float* __restrict__ out, int ld_out, float p, float inv_keep,
                   DropSeed seed, int site) {
  pdl_entry();
  const int total4 = R * C4, C = 4 * C4;
  const uint64_t sd = seed_value(seed);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += gridDim.x * blockDim.x) {
    const int r = q / C4, c = (q - r * C4) * 4;
    const float4 xv = x[q], m = *reinterpret_cast<const float4*>(mi + c), is = *reinterpret_cast<const float4*>(mi + C + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float4 y = make_float4((xv.x - m.x) * is.x * g.x + b.x, (xv.y - m.y) * is.y * g.y + b.y, (xv.z - m.z) * is.z * g.z + b.z,
                           (xv.w - m.w) * is.w * g.w + b.w);
    if (relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
    if (p > 0.f) y = drop4(y, sd, site, (uint64_t)q * 4, p, inv_keep);
    *reinterpret_cast<float4*>(out + (int64_t)r * ld_out + c) = y;
  }
}
// bn_finalize + bn_apply in one launch: every CTA derives mean / invstd of all C columns from the (all-reduced) double sums
// into shared memory (2C floats); CTA 0 also stores them for backward and updates the running statistics.
__global__ void __launch_bounds__(256)
bn_finalize_apply_v4_kernel(const double* __restrict__ sums, double n, float eps, int training, float momentum,
                            float* __restrict__ rmean, float* __restrict__ rvar, int64_t* __restrict__ tracked,
                            float* __restrict__ mi_out, const float4* __restrict__ x, int R, int C4,
                            const float* __restrict__ gamma, const float* __restrict__ beta, int relu, float* __restrict__ out,
                            int ld_out, float p, float inv_keep, DropSeed seed, int site) {
  pdl_entry();
  extern __shared__ __align__(16) float mis[];   // [2C]
  const int C = 4 * C4;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, inv;
    if (training) {
      const double m = sums[c] / n;
      double var = sums[C + c] / n - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      inv = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.x == 0 && rmean) {
        const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
      }
    } else {
      mean = rmean[c];
      inv = 1.0f / sqrtf(rvar[c] + eps);
    }
    mis[c] = mean;
    mis[C + c] = inv;
    if (blockIdx.x == 0) { mi_out[c] = mean; mi_out[C + c] = inv; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && training && tracked) *tracked += 1;
  __syncthreads();
  const int total4 = R * C4;
  const uint64_t sd = seed_value(seed);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += gridDim.x * blockDim.x) {
    const int r = q / C4, c = (q - r * C4) * 4;
    const float4 xv = x[q], m = *reinterpret_cast<const float4*>(mis + c), is = *reinterpret_cast<const float4*>(mis + C + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float4 y = make_float4((xv.x - m.x) * is.x * g.x + b.x, (xv.y - m.y) * is.y * g.y + b.y, (xv.z - m.z) * is.z * g.z + b.z,
                           (xv.w - m.w) * is.w * g.w + b.w);
    if (relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
    if (p > 0.f) y = drop4(y, sd, site, (uint64_t)q * 4, p, inv_keep);
    *reinterpret_cast<float4*>(out + (int64_t)r * ld_out + c) = y;
  }
}
// returns MVF_OK when the fused launch was taken, MVF_ERR_UNSUPPORTED (no error string) when the caller must use the two kernels
int bn_finalize_apply(const double* sums, int C, double n_global, float eps, int training, float momentum, float* rmean,
                      float* rvar, int64_t* tracked, float* mi, int dtype_out, const float* x, int64_t R,
                      const float* gamma, const float* beta, int relu, void* out, int64_t ld_out, float p, DropSeed seed, int site,
                      cudaStream_t st) {
  const int64_t total = R * C;
  if (!(dtype_out == MVF_F32 && C <= 4096 && total > 0 && vec4_ok(total, C, x, out, mi) && (ld_out & 3) == 0 && ld_out < (1ll << 31) &&
        ((((uintptr_t)gamma) | ((uintptr_t)beta)) & 15) == 0))
    return MVF_ERR_UNSUPPORTED;
  launch_k(bn_finalize_apply_v4_kernel, grid_for(total / 4, 1184), 256, (size_t)2 * C * sizeof(float), st, sums, n_global, eps, training,
           momentum, rmean, rvar, tracked, mi, (const float4*)x, (int)R, C / 4, gamma, beta, relu, (float*)out, (int)ld_out, p,
           p > 0.f ? 1.f / (1.f - p) : 1.f, seed, site);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

int bn_apply(int dtype_out, const float* x, int64_t R, int C, const float* mi, const float* gamma, const float* beta,
             int relu, void* out, int64_t ld_out, float p, DropSeed seed, int site, cudaStream_t st) {
  int64_t total = R * C;
  if (dtype_out == MVF_F32 && vec4_ok(total, C, x, out, mi) && (ld_out & 3) == 0 && ld_out < (1ll << 31) &&
      ((((uintptr_t)gamma) | ((uintptr_t)beta)) & 15) == 0) {
    launch_k(bn_apply_v4_kernel, grid_for(total / 4), 256, 0, st, (const float4*)x, (int)R, C / 4, mi, gamma, beta, relu, (float*)out,
             (int)ld_out, p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed, site);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  }
  int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  float ik = p > 0.f ? 1.f / (1.f - p) : 1.f;
  if (dtype_out == MVF_BF16)
    launch_k(bn_apply_kernel<bf16>, grid, 256, 0, st, x, R, C, mi, gamma, beta, relu, (bf16*)out, ld_out, p, ik, seed, site);
  else
    launch_k(bn_apply_kernel<float>, grid, 256, 0, st, x, R, C, mi, gamma, beta, relu, (float*)out, ld_out, p, ik, seed, site);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// dy = d_out * drop * [y > 0]; bsums[0:C] += sum dy; bsums[C:2C] += sum dy*xhat; dgamma/dbeta += local sums
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(const float* __restrict__ d_out, int64_t ld_d, const float* __restrict__ x, int64_t R, int C,
                    const float* __restrict__ mi, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int relu, float p, float inv_keep, DropSeed seed, int site, double* __restrict__ bsums,
                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_entry();
  __shared__ double sh[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    const float m = mi[c], is = mi[C + c], g = gamma[c], b = beta[c];
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < R; r += (int64_t)gridDim.y * 8) {
      float xh = (x[r * C + c] - m) * is;
      float dy = d_out[r * ld_d + c];
      if (p > 0.f) dy *= drop_scale(seed, site, (uint64_t)(r * C + c), p, inv_keep);
      if (relu && !(xh * g + b > 0.f)) dy = 0.f;
      s1 += (double)dy;
      s2 += (double)dy * (double)xh;
    }
  }
  sh[0][ty][tx] = s1;
  sh[1][ty][tx] = s2;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int j = 1; j < 8; ++j) { s1 += sh[0][j][tx]; s2 += sh[1][j][tx]; }
    atomicAdd(bsums + c, s1);
    atomicAdd(bsums + C + c, s2);
    atomicAdd(dbeta + c, (float)s1);
    atomicAdd(dgamma + c, (float)s2);
  }
}
int bn_bwd_stats(const float* d_out, int64_t ld_d, const float* x, int64_t R, int C, const float* mi,
                 const float* gamma, const float* beta, int relu, float p, DropSeed seed, int site, double* bsums,
                 float* dgamma, float* dbeta, cudaStream_t st) {
  int gy = cdiv(R, 8 * 16);
  if (gy > 64) gy = 64;
  if (gy < 1) gy = 1;
  float ik = p > 0.f ? 1.f / (1.f - p) : 1.f;
  launch_k(bn_bwd_stats_kernel, dim3(cdiv(C, 32), gy), 256, 0, st, d_out, ld_d, x, R, C, mi, gamma, beta, relu, p, ik, seed,
                                                             site, bsums, dgamma, dbeta);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// dx = gamma*invstd*(dy - sum(dy)/n - xhat*sum(dy*xhat)/n), n and sums global (torch SyncBatchNorm backward)
template <typename TO>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ d_out, int64_t ld_d, const float* __restrict__ x,
                                    int64_t R, int C, const float* __restrict__ mi, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, int relu, float p, float inv_keep, DropSeed seed,
                                    int site, const double* __restrict__ bsums, double n, TO* __restrict__ dx,
                                    int64_t ld_dx) {
  pdl_entry();
  int64_t total = R * C;
  const float inv_n = (float)(1.0 / n);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t r = i / C;
    const float m = mi[c], is = mi[C + c], g = gamma[c], b = beta[c];
    float xh = (x[i] - m) * is;
    float dy = d_out[r * ld_d + c];
    if (p > 0.f) dy *= drop_scale(seed, site, (uint64_t)i, p, inv_keep);
    if (relu && !(xh * g + b > 0.f)) dy = 0.f;
    float s1 = (float)bsums[c], s2 = (float)bsums[C + c];
    dx[r * ld_dx + c] = from_f<TO>(g * is * (dy - s1 * inv_n - xh * s2 * inv_n));
  }
}
__global__ void __launch_bounds__(256)
bn_bwd_apply_v4_kernel(const float* __restrict__ d_out, int ld_d, const float4* __restrict__ x, int R, int C4,
                       const float* __restrict__ mi, const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                       float p, float inv_keep, DropSeed seed, int site, const double* __restrict__ bsums, double n,
                       float* __restrict__ dx, int ld_dx) {
  pdl_entry();
  const int total4 = R * C4, C = 4 * C4;
  const float inv_n = (float)(1.0 / n);
  const uint64_t sd = seed_value(seed);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += gridDim.x * blockDim.x) {
    const int r = q / C4, c = (q - r * C4) * 4;
    const float4 xv = x[q], m = *reinterpret_cast<const float4*>(mi + c), is = *reinterpret_cast<const float4*>(mi + C + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float4 dy = *reinterpret_cast<const float4*>(d_out + (int64_t)r * ld_d + c);
    if (p > 0.f) dy = drop4(dy, sd, site, (uint64_t)q * 4, p, inv_keep);
    const float xh[4] = {(xv.x - m.x) * is.x, (xv.y - m.y) * is.y, (xv.z - m.z) * is.z, (xv.w - m.w) * is.w};
    const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w}, ii[4] = {is.x, is.y, is.z, is.w};
    float d[4] = {dy.x, dy.y, dy.z, dy.w}, o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (relu && !(xh[k] * gg[k] + bb[k] > 0.f)) d[k] = 0.f;
      const float s1 = (float)bsums[c + k], s2 = (float)bsums[C + c + k];
      o[k] = gg[k] * ii[k] * (d[k] - s1 * inv_n - xh[k] * s2 * inv_n);
    }
    *reinterpret_cast<float4*>(dx + (int64_t)r * ld_dx + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}
int bn_bwd_apply(int dtype_out, const float* d_out, int64_t ld_d, const float* x, int64_t R, int C, const float* mi,
                 const float* gamma, const float* beta, int relu, float p, DropSeed seed, int site, const double* bsums,
                 double n_global, void* dx, int64_t ld_dx, cudaStream_t st) {
  int64_t total = R * C;
  if (dtype_out == MVF_F32 && vec4_ok(total, C, x, dx, mi) && ((ld_d | ld_dx) & 3) == 0 && ld_d < (1ll << 31) && ld_dx < (1ll << 31) &&
      ((((uintptr_t)gamma) | ((uintptr_t)beta) | ((uintptr_t)d_out)) & 15) == 0) {
    launch_k(bn_bwd_apply_v4_kernel, grid_for(total / 4), 256, 0, st, d_out, (int)ld_d, (const float4*)x, (int)R, C / 4, mi, gamma, beta,
             relu, p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed, site, bsums, n_global, (float*)dx, (int)ld_dx);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  }
  int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  float ik = p > 0.f ? 1.f / (1.f - p) : 1.f;
  if (dtype_out == MVF_BF16)
    launch_k(bn_bwd_apply_kernel<bf16>, grid, 256, 0, st, d_out, ld_d, x, R, C, mi, gamma, beta, relu, p, ik, seed, site,
                                                    bsums, n_global, (bf16*)dx, ld_dx);
  else
    launch_k(bn_bwd_apply_kernel<float>, grid, 256, 0, st, d_out, ld_d, x, R, C, mi, gamma, beta, relu, p, ik, seed, site,
                                                     bsums, n_global, (float*)dx, ld_dx);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// =============================================== misc ================================================
template <typename TI>
__global__ void __launch_bounds__(256) colsum_kernel(const TI* __restrict__ X, int64_t R, int C, int64_t ld,
                                                     float* __restrict__ out) {
  pdl_entry();
  __shared__ float sh[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < C)
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < R; r += (int64_t)gridDim.y * 8) s += to_f<TI>(X[r * ld + c]);
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int j = 1; j < 8; ++j) s += sh[j][tx];
    atomicAdd(out + c, s);
  }
}
// Batched variant: every bias gradient of one backward call in ONE launch (blockIdx.z = matrix).  The dY buffers all stay
// alive until the side stream is joined, so the column sums can be deferred to the end instead of costing one launch each.
struct ColsumTable {
  int n;
  ColsumEntry e[COLSUM_MAX];
};
__global__ void __launch_bounds__(256) colsum_batched_kernel(const ColsumTable tab) {
  pdl_entry();
  const ColsumEntry& en = tab.e[blockIdx.z];
  __shared__ float sh[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (blockIdx.x * 32 >= en.C) return;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < en.C)
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < en.R; r += (int64_t)gridDim.y * 8) s += en.X[r * en.ld + c];
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < en.C) {
    for (int j = 1; j < 8; ++j) s += sh[j][tx];
    atomicAdd(en.out + c, s);
  }
}
int colsum_batched(const ColsumEntry* entries, int n, cudaStream_t st) {
  for (int base = 0; base < n; base += COLSUM_MAX) {
    ColsumTable tab;
    tab.n = n - base < COLSUM_MAX ? n - base : COLSUM_MAX;
    int maxC = 1;
    int64_t maxR = 1;
    for (int i = 0; i < tab.n; ++i) {
      tab.e[i] = entries[base + i];
      if (tab.e[i].C > maxC) maxC = tab.e[i].C;
      if (tab.e[i].R > maxR) maxR = tab.e[i].R;
    }
    int gy = cdiv(maxR, 8 * 32);
    gy = gy > 32 ? 32 : (gy < 1 ? 1 : gy);
    launch_k(colsum_batched_kernel, dim3(cdiv(maxC, 32), gy, tab.n), 256, 0, st, tab);
    MVF_CHECK_LAUNCH();
  }
  return MVF_OK;
}

int colsum(int dtype_in, const void* X, int64_t R, int C, int64_t ld, float* out, cudaStream_t st) {
  int gy = cdiv(R, 8 * 32);
  if (gy > 128) gy = 128;
  if (gy < 1) gy = 1;
  dim3 grid(cdiv(C, 32), gy);
  if (dtype_in == MVF_BF16) launch_k(colsum_kernel<bf16>, grid, 256, 0, st, (const bf16*)X, R, C, ld, out);
  else launch_k(colsum_kernel<float>, grid, 256, 0, st, (const float*)X, R, C, ld, out);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

template <typename TO>
__global__ void dropout_cast_kernel(const float* __restrict__ in, TO* __restrict__ out, int64_t rows, int cols,
                                    int64_t ld_out, float p, float inv_keep, DropSeed seed, int site) {
  pdl_entry();
  int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v = in[i];
    if (p > 0.f) v *= drop_scale(seed, site, (uint64_t)i, p, inv_keep);
    out[(i / cols) * ld_out + (i % cols)] = from_f<TO>(v);
  }
}
__global__ void __launch_bounds__(256)
dropout_cast_v4_kernel(const float4* __restrict__ in, float* __restrict__ out, int rows, int C4, int ld_out, float p,
                       float inv_keep, DropSeed seed, int site) {
  pdl_entry();
  const int total4 = rows * C4;
  const uint64_t sd = seed_value(seed);
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += gridDim.x * blockDim.x) {
    const int r = q / C4, c = (q - r * C4) * 4;
    float4 v = in[q];
    if (p > 0.f) v = drop4(v, sd, site, (uint64_t)q * 4, p, inv_keep);
    *reinterpret_cast<float4*>(out + (int64_t)r * ld_out + c) = v;
  }
}
int dropout_cast(int dtype_out, const float* in, void* out, int64_t rows, int cols, int64_t ld_out, float p,
                 DropSeed seed, int site, cudaStream_t st) {
  int64_t total = rows * cols;
  if (total == 0) return MVF_OK;
  if (dtype_out == MVF_F32 && vec4_ok(total, cols, in, out) && (ld_out & 3) == 0 && ld_out < (1ll << 31)) {
    launch_k(dropout_cast_v4_kernel, grid_for(total / 4), 256, 0, st, (const float4*)in, (float*)out, (int)rows, cols / 4, (int)ld_out,
             p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed, site);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  }
  int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  float ik = p > 0.f ? 1.f / (1.f - p) : 1.f;
  if (dtype_out == MVF_BF16)
    launch_k(dropout_cast_kernel<bf16>, grid, 256, 0, st, in, (bf16*)out, rows, cols, ld_out, p, ik, seed, site);
  else
    launch_k(dropout_cast_kernel<float>, grid, 256, 0, st, in, (float*)out, rows, cols, ld_out, p, ik, seed, site);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
int cast_f32(int dtype_out, const float* in, void* out, int64_t n, cudaStream_t st) {
  return dropout_cast(dtype_out, in, out, 1, (int)n, n, 0.f, 0, 0, st);
}

__global__ void dropout_mask_kernel(DropSeed seed, int site, int64_t total, float p, float inv_keep, float* out) {
  pdl_entry();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = p > 0.f ? drop_scale(seed, site, (uint64_t)i, p, inv_keep) : 1.f;
}
int dropout_mask_export(DropSeed seed, int site, int64_t rows, int64_t cols, float p, float* out, cudaStream_t st) {
  int64_t total = rows * cols;
  if (total == 0) return MVF_OK;
  int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  launch_k(dropout_mask_kernel, grid, 256, 0, st, seed, site, total, p, p > 0.f ? 1.f / (1.f - p) : 1.f, out);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// entity reduction over e of z[b, e*T+t, c] (mvformer.py:181-190)
template <typename TO>
__global__ void entity_reduce_fwd_kernel(const float* __restrict__ z, TO* __restrict__ y, int32_t* __restrict__ argmax,
                                         int BV, int T, int E, int H, int mode) {
  pdl_entry();
  int64_t total = (int64_t)BV * T * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % H);
    int64_t bt = i / H;
    int t = (int)(bt % T), b = (int)(bt / T);
    const float* zp = z + ((int64_t)b * E * T + t) * H + c;
    const int64_t es = (int64_t)T * H;
    float v;
    if (mode == MVF_FINAL_ONE) v = zp[0];
    else if (mode == MVF_FINAL_AVG) {
      v = 0.f;
      for (int e = 0; e < E; ++e) v += zp[e * es];
      v /= (float)E;
    } else {
      v = zp[0];
      int am = 0;
      for (int e = 1; e < E; ++e) {
        float w = zp[e * es];
        if (w > v) { v = w; am = e; }
      }
      if (argmax) argmax[i] = am;
    }
    y[i] = from_f<TO>(v);
  }
}
int entity_reduce_fwd(int dtype_out, const float* z, void* y, int32_t* argmax, int BV, int T, int E, int H, int mode,
                      cudaStream_t st) {
  int64_t total = (int64_t)BV * T * H;
  int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  if (dtype_out == MVF_BF16) launch_k(entity_reduce_fwd_kernel<bf16>, grid, 256, 0, st, z, (bf16*)y, argmax, BV, T, E, H, mode);
  else launch_k(entity_reduce_fwd_kernel<float>, grid, 256, 0, st, z, (float*)y, argmax, BV, T, E, H, mode);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
__global__ void entity_reduce_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ argmax,
                                         float* __restrict__ dz, int BV, int T, int E, int H, int mode) {
  pdl_entry();
  int64_t total = (int64_t)BV * E * T * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % H);
    int64_t row = i / H;
    int t = (int)(row % T);
    int e = (int)((row / T) % E);
    int b = (int)(row / ((int64_t)T * E));
    int64_t yi = ((int64_t)b * T + t) * H + c;
    float g = dy[yi];
    float v;
    if (mode == MVF_FINAL_ONE) v = e == 0 ? g : 0.f;
    else if (mode == MVF_FINAL_AVG) v = g / (float)E;
    else v = argmax[yi] == e ? g : 0.f;
    dz[i] = v;
  }
}
__global__ void __launch_bounds__(256)
entity_reduce_bwd_v4_kernel(const float4* __restrict__ dy, const int4* __restrict__ argmax, float4* __restrict__ dz, int BV, int T,
                            int E, int H4, int mode) {
  pdl_entry();
  const int total4 = BV * E * T * H4;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += gridDim.x * blockDim.x) {
    const int row = q / H4, c4 = q - row * H4;
    const int be = row / T, t = row - be * T;
    const int b = be / E, e = be - b * E;
    const int yi = (b * T + t) * H4 + c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mode == MVF_FINAL_ONE) {
      if (e == 0) v = dy[yi];
    } else if (mode == MVF_FINAL_AVG) {
      const float4 g = dy[yi];
      v = make_float4(g.x / (float)E, g.y / (float)E, g.z / (float)E, g.w / (float)E);
    } else {
      const float4 g = dy[yi];
      const int4 a = argmax[yi];
      v = make_float4(a.x == e ? g.x : 0.f, a.y == e ? g.y : 0.f, a.z == e ? g.z : 0.f, a.w == e ? g.w : 0.f);
    }
    dz[q] = v;
  }
}
int entity_reduce_bwd(const float* dy, const int32_t* argmax, float* dz, int BV, int T, int E, int H, int mode,
                      cudaStream_t st) {
  int64_t total = (int64_t)BV * E * T * H;
  if (vec4_ok(total, H, dy, dz, argmax)) {
    launch_k(entity_reduce_bwd_v4_kernel, grid_for(total / 4), 256, 0, st, (const float4*)dy, (const int4*)argmax, (float4*)dz, BV, T,
             E, H / 4, mode);
    MVF_CHECK_LAUNCH();
    return MVF_OK;
  }
  int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  launch_k(entity_reduce_bwd_kernel, grid, 256, 0, st, dy, argmax, dz, BV, T, E, H, mode);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// SMART_FINAL 'lin' (mvformer.py:191-195): zl[b*T+t, e*H+c] = z[b, e*T+t, c]
template <typename TO>
__global__ void entity_gather_lin_kernel(const float* __restrict__ z, TO* __restrict__ zl, int BV, int T, int E, int H) {
  pdl_entry();
  int64_t total = (int64_t)BV * E * T * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % H);
    int64_t row = i / H;
    int t = (int)(row % T);
    int e = (int)((row / T) % E);
    int b = (int)(row / ((int64_t)T * E));
    zl[(((int64_t)b * T + t) * E + e) * H + c] = from_f<TO>(z[i]);
  }
}
int entity_gather_lin(int dtype_out, const float* z, void* zl, int BV, int T, int E, int H, cudaStream_t st) {
  int64_t total = (int64_t)BV * E * T * H;
  int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  if (dtype_out == MVF_BF16) launch_k(entity_gather_lin_kernel<bf16>, grid, 256, 0, st, z, (bf16*)zl, BV, T, E, H);
  else launch_k(entity_gather_lin_kernel<float>, grid, 256, 0, st, z, (float*)zl, BV, T, E, H);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
__global__ void entity_scatter_lin_kernel(const float* __restrict__ dzl, float* __restrict__ dz, int BV, int T, int E,
                                          int H) {
  pdl_entry();
  int64_t total = (int64_t)BV * E * T * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % H);
    int64_t row = i / H;
    int t = (int)(row % T);
    int e = (int)((row / T) % E);
    int b = (int)(row / ((int64_t)T * E));
    dz[i] = dzl[(((int64_t)b * T + t) * E + e) * H + c];
  }
}
int entity_scatter_lin(const float* dzl, float* dz, int BV, int T, int E, int H, cudaStream_t st) {
  int64_t total = (int64_t)BV * E * T * H;
  int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  launch_k(entity_scatter_lin_kernel, grid, 256, 0, st, dzl, dz, BV, T, E, H);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

// F.normalize(x, dim=-1, eps=1e-12): y = x / max(||x||, eps)   (transformer.py:228)
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                         float* __restrict__ norm, int64_t rows, int D,
                                                         float* __restrict__ y2) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) { float v = x[row * D + c]; s += v * v; }
  s = warp_sum(s);
  float n = fmaxf(sqrtf(s), 1e-12f);
  for (int c = lane; c < D; c += 32) {
    const float v = x[row * D + c] / n;
    y[row * D + c] = v;
    if (y2) y2[row * D + c] = v;      // the caller's output tensor (the saved copy stays for backward)
  }
  if (lane == 0) norm[row] = n;
}
int l2norm_fwd(const float* x, float* y, float* norm, int64_t rows, int D, cudaStream_t st, float* y2) {
  launch_k(l2norm_fwd_kernel, cdiv(rows, 8), 256, 0, st, x, y, norm, rows, D, y2);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}
// dx = (dy - y * <y, dy>) / n   (n = clamped norm; for ||x|| <= eps the clamp is active: dx = dy / eps)
template <typename TO>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                         const float* __restrict__ norm, TO* __restrict__ dx,
                                                         int64_t rows, int D) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += y[row * D + c] * dy[row * D + c];
  s = warp_sum(s);
  const float n = norm[row];
  if (n <= 1e-12f) s = 0.f;
  for (int c = lane; c < D; c += 32) dx[row * D + c] = from_f<TO>((dy[row * D + c] - y[row * D + c] * s) / n);
}
int l2norm_bwd(const float* dy, const float* y, const float* norm, void* dx, int dtype_out, int64_t rows, int D,
               cudaStream_t st) {
  if (dtype_out == MVF_BF16) launch_k(l2norm_bwd_kernel<bf16>, cdiv(rows, 8), 256, 0, st, dy, y, norm, (bf16*)dx, rows, D);
  else launch_k(l2norm_bwd_kernel<float>, cdiv(rows, 8), 256, 0, st, dy, y, norm, (float*)dx, rows, D);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace mvf

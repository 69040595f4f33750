// Fused optimizer tail of the training step: loss-scale removal + global-norm gradient clipping
// (train.py:124-133, torch.nn.utils.clip_grad_norm_ with OPTIMIZER.GRAD_CLIP) + Adam / AdamW (utils/optimizer.py:60-73,
// betas (0.9, 0.999), weight decay on every group) over ALL trainable head tensors in two launches:
//   opt_sqnorm_kernel  per-CTA partial sums of g^2 (float64), table-driven over the parameter list
//   opt_adam_kernel    every CTA re-adds the partials (fixed order -> the same clip factor everywhere), then updates
// The step count and the learning rate live in device memory (the kernel advances the count), so both launches can sit
// at the end of the captured step graph.  Arithmetic follows torch.optim.Adam's single-tensor path:
//   g' = g * inv_scale * min(1, max_norm / (||g * inv_scale|| + 1e-6));  Adam: g' += wd * p;  AdamW: p *= 1 - lr * wd
//   m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include <math.h>

#include "kernels.cuh"

namespace mvf {

constexpr int OPT_CHUNK = 48;       // tensors per launch (the table travels as a kernel parameter)
constexpr int OPT_GX = 74;          // CTAs per tensor

struct OptEntry {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
};
struct OptTable {
  OptEntry e[OPT_CHUNK];
  int n;
};

__global__ void __launch_bounds__(256)
opt_sqnorm_kernel(const OptTable tab, double* __restrict__ partial, int part_base) {
  pdl_entry();
  __shared__ double red[8];
  const OptEntry& en = tab.e[blockIdx.y];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < en.n; i += (int64_t)gridDim.x * blockDim.x) {
    const double g = (double)en.g[i];
    acc += g * g;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    partial[part_base + blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
opt_adam_kernel(const OptTable tab, const double* __restrict__ partial, int nparts, const float* __restrict__ lr_dev,
                const int64_t* __restrict__ step_dev, double beta1d, double beta2d, float eps, float wd, int adamw, float max_norm,
                float inv_scale, float* __restrict__ norm_out) {
  pdl_entry();
  __shared__ double red[8];
  __shared__ float s_clip;
  // total norm: every CTA adds the same partials in the same order
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    const float norm = (float)sqrt(t) * inv_scale;
    float clip = 1.f;
    if (max_norm > 0.f) clip = fminf(1.f, max_norm / (norm + 1e-6f));
    s_clip = clip * inv_scale;
    if (blockIdx.x == 0 && blockIdx.y == 0 && norm_out) *norm_out = norm;
  }
  __syncthreads();
  const float gmul = s_clip;
  const float lr = *lr_dev;
  const double t = (double)(*step_dev + 1);                 // this step's count (the counter is advanced after the launch)
  // the hyper-parameters arrive as doubles: torch forms 1 - beta in Python floats (1 - 0.999f would be off by 1.3e-5)
  const float beta2 = (float)beta2d, omb1 = (float)(1.0 - beta1d), omb2 = (float)(1.0 - beta2d);
  const float bc1 = (float)(1.0 - pow(beta1d, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2d, t));
  const float step_size = lr / bc1;
  const OptEntry& en = tab.e[blockIdx.y];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < en.n; i += (int64_t)gridDim.x * blockDim.x) {
    float p = en.p[i];
    float g = en.g[i] * gmul;
    if (wd != 0.f) {
      if (adamw) p *= 1.f - lr * wd;
      else g = fmaf(wd, p, g);
    }
    const float m = en.m[i] + (g - en.m[i]) * omb1;                 // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float v = fmaf(en.v[i], beta2, omb2 * g * g);             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    en.m[i] = m;
    en.v[i] = v;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    en.p[i] = p - step_size * (m / denom);
  }
}

__global__ void opt_advance_kernel(int64_t* step_dev) {
  pdl_entry();
  if (threadIdx.x == 0 && blockIdx.x == 0) *step_dev += 1;
}

size_t opt_ws_bytes(int n_tensors) { return sizeof(double) * (size_t)n_tensors * OPT_GX; }

int opt_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* m, float* const* v,
                  const int64_t* numel, const float* lr_dev, int64_t* step_dev, double beta1, double beta2, float eps, float wd,
                  int adamw, float max_norm, float inv_scale, float* norm_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  MVF_REQUIRE(n_tensors >= 0 && params && grads && m && v && numel && lr_dev && step_dev && ws, MVF_ERR_BAD_ARG,
              "opt_adam_step: null pointer");
  MVF_REQUIRE(ws_bytes >= opt_ws_bytes(n_tensors), MVF_ERR_WORKSPACE, "opt_adam_step: workspace %zu < %zu bytes", ws_bytes,
              opt_ws_bytes(n_tensors));
  if (n_tensors == 0) return MVF_OK;
  double* partial = (double*)ws;
  auto fill = [&](OptTable& tab, int base) {
    tab.n = (n_tensors - base < OPT_CHUNK) ? n_tensors - base : OPT_CHUNK;
    for (int i = 0; i < tab.n; ++i) tab.e[i] = OptEntry{params[base + i], grads[base + i], m[base + i], v[base + i], numel[base + i]};
  };
  for (int base = 0; base < n_tensors; base += OPT_CHUNK) {
    OptTable tab;
    fill(tab, base);
    launch_k(opt_sqnorm_kernel, dim3(OPT_GX, tab.n), 256, 0, st, tab, partial, base * OPT_GX);
    MVF_CHECK_LAUNCH();
  }
  for (int base = 0; base < n_tensors; base += OPT_CHUNK) {
    OptTable tab;
    fill(tab, base);
    launch_k(opt_adam_kernel, dim3(OPT_GX, tab.n), 256, 0, st, tab, (const double*)partial, n_tensors * OPT_GX, lr_dev,
             (const int64_t*)step_dev, beta1, beta2, eps, wd, adamw, max_norm, inv_scale, norm_out);
    MVF_CHECK_LAUNCH();
  }
  launch_k(opt_advance_kernel, 1, 32, 0, st, step_dev);
  MVF_CHECK_LAUNCH();
  return MVF_OK;
}

}  // namespace mvf

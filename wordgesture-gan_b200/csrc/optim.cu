// Fused gradient-clip + Adam over a module's flat parameter / gradient / moment buffers.
// Replaces torch.nn.utils.clip_grad_norm_(module.parameters(), max_norm) followed by
// torch.optim.Adam.step() (src/shared/utils.py:87-88,108-109,132-135; Adam built at
// src/gan/trainer.py:60-79: lr 2e-4 (scheduler-mutated), betas (0.5, 0.999), eps 1e-8, no weight decay).
//   total_norm = ||g||_2 ; coef = min(1, max_norm / (total_norm + 1e-6)) ; g *= coef        (clip_grad_norm_)
//   m += (g - m)(1-b1) ; v = b2 v + (1-b2) g^2 ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)   (Adam)
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kNormBlocks = 128;

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  __shared__ float red[33];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        const float* __restrict__ partial, int npartial,
                                                        float max_norm, float step_size, float beta1, float beta2,
                                                        float bc2_sqrt, float eps, float* __restrict__ norm_out) {
  __shared__ float s_coef;
  if (threadIdx.x == 0) {
    float coef = 1.f;
    if (max_norm > 0.f) {
      float tot = 0.f;
      for (int i = 0; i < npartial; ++i) tot += partial[i];  // same order in every block: deterministic
      const float norm = sqrtf(tot);
      coef = fminf(max_norm / (norm + 1e-6f), 1.f);
      if (norm_out && blockIdx.x == 0) norm_out[0] = norm;
    }
    s_coef = coef;
  }
  __syncthreads();
  const float coef = s_coef;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    g[i] = gi;
    float mi = m[i];
    mi = mi + (gi - mi) * (1.f - beta1);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// Device-state variant (CUDA-graph capturable: nothing step-dependent is baked into the launch).  The step counter
// lives on the device and is advanced by the first kernel; bias corrections are evaluated in double on the device.
__global__ void __launch_bounds__(256) sumsq_step_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial,
                                                         int* __restrict__ step_dev, int do_norm) {
  __shared__ float red[33];
  if (blockIdx.x == 0 && threadIdx.x == 0) step_dev[0] += 1;
  if (!do_norm) return;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) clip_adam_dev_kernel(float* __restrict__ p, float* __restrict__ g,
                                                            float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                            const float* __restrict__ partial, int npartial,
                                                            float max_norm, const float* __restrict__ lr_dev, float beta1,
                                                            float beta2, float eps, const int* __restrict__ step_dev,
                                                            float* __restrict__ norm_out) {
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    float coef = 1.f;
    if (max_norm > 0.f) {
      float tot = 0.f;
      for (int i = 0; i < npartial; ++i) tot += partial[i];
      const float norm = sqrtf(tot);
      coef = fminf(max_norm / (norm + 1e-6f), 1.f);
      if (norm_out && blockIdx.x == 0) norm_out[0] = norm;
    }
    s_coef = coef;
    const double step = (double)step_dev[0];
    const double bc1 = 1.0 - pow((double)beta1, step);
    const double bc2 = 1.0 - pow((double)beta2, step);
    s_step_size = (float)((double)lr_dev[0] / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    g[i] = gi;
    float mi = m[i];
    mi = mi + (gi - mi) * (1.f - beta1);
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

}  // namespace

extern "C" int wgg_clip_adam_dev(wgg_ctx* ctx, float* p, float* g, float* m, float* v, int64_t n, const float* lr_dev,
                                 float beta1, float beta2, float eps, int* step_dev, float max_norm,
                                 float* grad_norm_out, float* ws, void* stream) {
  if (!ctx || n <= 0 || !lr_dev || !step_dev || !ws) return WGG_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int nb = (int)cdiv64(n, 256 * 8);
  if (nb > kNormBlocks) nb = kNormBlocks;
  if (nb < 1) nb = 1;
  sumsq_step_kernel<<<nb, 256, 0, st>>>(g, n, ws, step_dev, max_norm > 0.f ? 1 : 0);
  WGG_CHECK_LAUNCH(ctx, "sumsq_step_kernel");
  int ab = (int)cdiv64(n, 256 * 4);
  if (ab > 148 * 4) ab = 148 * 4;
  if (ab < 1) ab = 1;
  clip_adam_dev_kernel<<<ab, 256, 0, st>>>(p, g, m, v, n, ws, nb, max_norm, lr_dev, beta1, beta2, eps, step_dev,
                                           grad_norm_out);
  WGG_CHECK_LAUNCH(ctx, "clip_adam_dev_kernel");
  return WGG_OK;
}

extern "C" int64_t wgg_clip_adam_workspace_floats(void) { return kNormBlocks; }

extern "C" int wgg_clip_adam(wgg_ctx* ctx, float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, int64_t step, float max_norm, float* grad_norm_out, float* ws,
                             void* stream) {
  if (!ctx || n <= 0 || step < 1) return WGG_EINVAL;
  if (max_norm > 0.f && !ws) return wgg_fail(ctx, WGG_EWORKSPACE, "clip_adam: workspace missing%s");
  cudaStream_t st = (cudaStream_t)stream;
  int nb = (int)cdiv64(n, 256 * 8);
  if (nb > kNormBlocks) nb = kNormBlocks;
  if (nb < 1) nb = 1;
  if (max_norm > 0.f) {
    sumsq_kernel<<<nb, 256, 0, st>>>(g, n, ws);
    WGG_CHECK_LAUNCH(ctx, "sumsq_kernel");
  }
  // bias corrections in double on the host, as torch.optim.Adam's scalar path does
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  int ab = (int)cdiv64(n, 256 * 4);
  if (ab > 148 * 4) ab = 148 * 4;
  if (ab < 1) ab = 1;
  clip_adam_kernel<<<ab, 256, 0, st>>>(p, g, m, v, n, ws, nb, max_norm, step_size, beta1, beta2, bc2_sqrt, eps,
                                       grad_norm_out);
  WGG_CHECK_LAUNCH(ctx, "clip_adam_kernel");
  return WGG_OK;
}

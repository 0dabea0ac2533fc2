// Small shared elementwise kernels.
#include "common.cuh"

namespace {
__global__ void leaky_bwd_kernel(const float* __restrict__ y, float* __restrict__ d, const float* __restrict__ add,
                                 int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = d[i];
    if (add) g += __ldg(add + i);
    d[i] = __ldg(y + i) > 0.f ? g : kLeak * g;
  }
}
__global__ void fill_kernel(float* __restrict__ x, float v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
}  // namespace

int leaky_bwd_launch(wgg_ctx* ctx, const float* y, float* d, const float* add, int64_t n, cudaStream_t st) {
  if (n <= 0) return WGG_OK;
  leaky_bwd_kernel<<<ew_blocks(n), 256, 0, st>>>(y, d, add, n);
  WGG_CHECK_LAUNCH(ctx, "leaky_bwd_kernel");
  return WGG_OK;
}

int fill_launch(wgg_ctx* ctx, float* x, float v, int64_t n, cudaStream_t st) {
  if (n <= 0) return WGG_OK;
  fill_kernel<<<ew_blocks(n), 256, 0, st>>>(x, v, n);
  WGG_CHECK_LAUNCH(ctx, "fill_kernel");
  return WGG_OK;
}

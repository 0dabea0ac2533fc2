// Loss kernels: Wasserstein means, L1 (reconstruction / latent), feature matching, KL.
// Replaces src/gan/losses.py: WassersteinLoss :43,:58; FeatureMatchingLoss :86-93; ReconstructionLoss :120;
// LatentEncodingLoss :147; KLDivergenceLoss :174-175.  Scalars stay on the device (no .item() sync);
// reductions are two-stage and deterministic (fixed block partition, fixed summation order).
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int kMaxSeg = WGG_MAX_HIDDEN_LAYERS + 2;

// weighted segmented reduction:  sum_k w[k] * sum_{i in seg k} f(a[off_k + i], b[off_k + i])
struct SegTable {
  int n;
  int64_t off[kMaxSeg];
  int64_t end[kMaxSeg];  // cumulative element counts (virtual index space)
  float w[kMaxSeg];
};

enum { OP_SUM = 0, OP_ABSDIFF = 1 };

template <int OP>
__global__ void __launch_bounds__(256) seg_reduce_kernel(SegTable tb, const float* __restrict__ a,
                                                         const float* __restrict__ b, float* __restrict__ partial) {
  __shared__ float red[33];
  const int64_t total = tb.end[tb.n - 1];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= tb.end[k]) ++k;
    const int64_t j = tb.off[k] + (i - (k ? tb.end[k - 1] : 0));
    float v;
    if (OP == OP_SUM) v = __ldg(a + j);
    else v = fabsf(__ldg(a + j) - __ldg(b + j));
    acc = fmaf(tb.w[k], v, acc);
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// float4 variant: every segment offset / length is a multiple of 4 and the bases are 16-byte aligned
template <int OP>
__global__ void __launch_bounds__(256) seg_reduce4_kernel(SegTable tb, const float4* __restrict__ a,
                                                          const float4* __restrict__ b, float* __restrict__ partial) {
  __shared__ float red[33];
  const int64_t total = tb.end[tb.n - 1] >> 2;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= (tb.end[k] >> 2)) ++k;
    const int64_t j = (tb.off[k] >> 2) + (i - (k ? (tb.end[k - 1] >> 2) : 0));
    const float4 x = __ldg(a + j);
    float v;
    if (OP == OP_SUM) v = (x.x + x.y) + (x.z + x.w);
    else {
      const float4 y = __ldg(b + j);
      v = (fabsf(x.x - y.x) + fabsf(x.y - y.y)) + (fabsf(x.z - y.z) + fabsf(x.w - y.w));
    }
    acc = fmaf(tb.w[k], v, acc);
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void finalize_kernel(const float* __restrict__ partial, int n, float scale, int accumulate,
                                float* __restrict__ out) {
  __shared__ float red[33];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + scale * s;
}

// d[off_k + i] (+)= g * w[k] * sign(a - b)
__global__ void __launch_bounds__(256) seg_sign_kernel(SegTable tb, const float* __restrict__ a,
                                                       const float* __restrict__ b, const float* __restrict__ g,
                                                       int accumulate, float* __restrict__ d) {
  const int64_t total = tb.end[tb.n - 1];
  const float gg = g ? __ldg(g) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= tb.end[k]) ++k;
    const int64_t j = tb.off[k] + (i - (k ? tb.end[k - 1] : 0));
    const float df = __ldg(a + j) - __ldg(b + j);
    const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
    const float v = gg * tb.w[k] * sgn;
    d[j] = accumulate ? d[j] + v : v;
  }
}

__global__ void mean_bwd_kernel(const float* __restrict__ g, float coef, int64_t n, float* __restrict__ dx) {
  const float v = (g ? __ldg(g) : 1.f) * coef;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dx[i] = v;
}

// kld_b = -0.5 * sum_j (1 + lv - mu^2 - exp(lv)); partial sums of kld over b
__global__ void __launch_bounds__(256) kl_kernel(const float* __restrict__ mu, const float* __restrict__ lv, int64_t n,
                                                 float* __restrict__ partial) {
  __shared__ float red[33];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = __ldg(mu + i), l = __ldg(lv + i);
    acc += 1.f + l - m * m - expf(l);
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = -0.5f * s;
}

__global__ void kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ g,
                              float coef, int64_t n, float* __restrict__ dmu, float* __restrict__ dlv) {
  const float gg = (g ? __ldg(g) : 1.f) * coef;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    dmu[i] = gg * __ldg(mu + i);
    dlv[i] = gg * 0.5f * (expf(__ldg(lv + i)) - 1.f);
  }
}

int red_blocks(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > kRedBlocks) g = kRedBlocks;
  if (g < 1) g = 1;
  return (int)g;
}

float* next_partial(wgg_ctx* ctx) { return ctx->red_scratch ? wgg_next_partial(ctx) : nullptr; }

template <int OP>
int seg_reduce(wgg_ctx* ctx, const SegTable& tb, const float* a, const float* b, float scale, int accumulate,
               float* out, cudaStream_t st) {
  float* partial = next_partial(ctx);
  if (!partial) return wgg_fail(ctx, WGG_ECUDA, "loss: cannot allocate reduction scratch%s");
  bool vec = (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
  for (int k = 0; k < tb.n; ++k) vec = vec && (tb.off[k] % 4 == 0) && (tb.end[k] % 4 == 0);
  const int nb = red_blocks(vec ? tb.end[tb.n - 1] / 4 : tb.end[tb.n - 1]);
  if (vec) seg_reduce4_kernel<OP><<<nb, 256, 0, st>>>(tb, reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), partial);
  else seg_reduce_kernel<OP><<<nb, 256, 0, st>>>(tb, a, b, partial);
  WGG_CHECK_LAUNCH(ctx, "seg_reduce_kernel");
  finalize_kernel<<<1, 256, 0, st>>>(partial, nb, scale, accumulate, out);
  WGG_CHECK_LAUNCH(ctx, "finalize_kernel");
  return WGG_OK;
}

SegTable one_seg(int64_t n, float w) {
  SegTable tb;
  tb.n = 1; tb.off[0] = 0; tb.end[0] = n; tb.w[0] = w;
  return tb;
}

int fm_table(const wgg_model_cfg* cfg, int64_t B, SegTable* tb) {
  FeatTable ft;
  if (disc_feature_table(cfg, B, &ft) != WGG_OK) return WGG_EINVAL;
  tb->n = ft.n;
  int64_t cum = 0;
  for (int k = 0; k < ft.n; ++k) {
    tb->off[k] = ft.off[k];
    cum += ft.count[k];
    tb->end[k] = cum;
    // losses.py:88-93:  (1/K) * sum_k  mean|f-r| / n_k   with mean over B*n_k elements
    tb->w[k] = (float)(1.0 / ((double)ft.count[k] * (double)ft.width[k] * (double)ft.n));
  }
  return WGG_OK;
}

}  // namespace

extern "C" int wgg_mean(wgg_ctx* ctx, const float* x, int64_t n, float scale, int accumulate, float* out,
                        void* stream) {
  if (!ctx || n <= 0) return WGG_EINVAL;
  return seg_reduce<OP_SUM>(ctx, one_seg(n, 1.f), x, nullptr, (float)((double)scale / (double)n), accumulate, out,
                            (cudaStream_t)stream);
}

extern "C" int wgg_mean_backward(wgg_ctx* ctx, const float* g, float scale, int64_t n, float* dx, void* stream) {
  if (!ctx || n <= 0) return WGG_EINVAL;
  mean_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(g, (float)((double)scale / (double)n), n, dx);
  WGG_CHECK_LAUNCH(ctx, "mean_bwd_kernel");
  return WGG_OK;
}

extern "C" int wgg_l1_mean(wgg_ctx* ctx, const float* a, const float* b, int64_t n, float scale, int accumulate,
                           float* out, void* stream) {
  if (!ctx || n <= 0) return WGG_EINVAL;
  return seg_reduce<OP_ABSDIFF>(ctx, one_seg(n, 1.f), a, b, (float)((double)scale / (double)n), accumulate, out,
                                (cudaStream_t)stream);
}

extern "C" int wgg_l1_mean_backward(wgg_ctx* ctx, const float* a, const float* b, const float* g, float scale,
                                    int64_t n, int accumulate, float* da, void* stream) {
  if (!ctx || n <= 0) return WGG_EINVAL;
  seg_sign_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(one_seg(n, (float)((double)scale / (double)n)), a, b,
                                                                  g, accumulate, da);
  WGG_CHECK_LAUNCH(ctx, "seg_sign_kernel");
  return WGG_OK;
}

extern "C" int wgg_feature_matching(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* real_stash,
                                    const float* fake_stash, int64_t B, float scale, int accumulate, float* out,
                                    void* stream) {
  if (!ctx || B <= 0) return WGG_EINVAL;
  SegTable tb;
  if (fm_table(cfg, B, &tb) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "feature_matching: bad config%s");
  return seg_reduce<OP_ABSDIFF>(ctx, tb, fake_stash, real_stash, scale, accumulate, out, (cudaStream_t)stream);
}

extern "C" int wgg_feature_matching_backward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* real_stash,
                                             const float* fake_stash, const float* g, float scale, int64_t B,
                                             float* dfeat, void* stream) {
  if (!ctx || B <= 0) return WGG_EINVAL;
  SegTable tb;
  if (fm_table(cfg, B, &tb) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "feature_matching: bad config%s");
  for (int k = 0; k < tb.n; ++k) tb.w[k] *= scale;
  seg_sign_kernel<<<ew_blocks(tb.end[tb.n - 1]), 256, 0, (cudaStream_t)stream>>>(tb, fake_stash, real_stash, g, 0,
                                                                                 dfeat);
  WGG_CHECK_LAUNCH(ctx, "seg_sign_kernel");
  return WGG_OK;
}

int wgg_loss_finalize(wgg_ctx* ctx, const float* partial, int n, float scale, int accumulate, float* out, cudaStream_t st) {
  finalize_kernel<<<1, 256, 0, st>>>(partial, n, scale, accumulate, out);
  WGG_CHECK_LAUNCH(ctx, "finalize_kernel");
  return WGG_OK;
}

extern "C" int wgg_kl(wgg_ctx* ctx, const float* mu, const float* log_var, int64_t B, int32_t Z, float scale,
                      int accumulate, float* out, void* stream) {
  if (!ctx || B <= 0) return WGG_EINVAL;
  float* partial = next_partial(ctx);
  if (!partial) return wgg_fail(ctx, WGG_ECUDA, "loss: cannot allocate reduction scratch%s");
  const int nb = red_blocks(B * Z);
  kl_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(mu, log_var, B * Z, partial);
  WGG_CHECK_LAUNCH(ctx, "kl_kernel");
  finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partial, nb, (float)((double)scale / (double)B), accumulate, out);
  WGG_CHECK_LAUNCH(ctx, "finalize_kernel");
  return WGG_OK;
}

extern "C" int wgg_kl_backward(wgg_ctx* ctx, const float* mu, const float* log_var, const float* g, float scale,
                               int64_t B, int32_t Z, float* dmu, float* dlog_var, void* stream) {
  if (!ctx || B <= 0) return WGG_EINVAL;
  kl_bwd_kernel<<<ew_blocks(B * Z), 256, 0, (cudaStream_t)stream>>>(mu, log_var, g, (float)((double)scale / (double)B),
                                                                    B * Z, dmu, dlog_var);
  WGG_CHECK_LAUNCH(ctx, "kl_bwd_kernel");
  return WGG_OK;
}

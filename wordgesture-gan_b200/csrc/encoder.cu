// VariationalEncoder forward + reparameterisation and backward.
// Replaces VariationalEncoder.forward / reparameterize (src/gan/models.py:52-86): flatten :64,
// [Linear + LeakyReLU(0.2)] x n :67, fc_mu / fc_log_var :70-71, z = mu + eps * exp(0.5 log_var) :84-86.
// eps is drawn by the caller with torch.randn (the RNG stream stays PyTorch's).
// The latent head (both heads + reparameterisation + KL partial sums, losses.py:174-175) is one fused kernel.
#include "common.cuh"

namespace {

struct EncLayout {
  int n;                                   // hidden layers
  int dims[WGG_MAX_HIDDEN_LAYERS + 1];     // dims[0] = T*C input, dims[i+1] = hidden i
  int Z;
  int64_t off_w[WGG_MAX_HIDDEN_LAYERS], off_b[WGG_MAX_HIDDEN_LAYERS];
  int64_t off_wmu, off_bmu, off_wlv, off_blv, total;
  int64_t act_off[WGG_MAX_HIDDEN_LAYERS + 1];  // per-sample offsets into the stash, by layer (x B)
  int64_t act_width;
};

int enc_layout(const wgg_model_cfg* c, EncLayout* e) {
  if (!c || c->n_enc_hidden < 1 || c->n_enc_hidden > WGG_MAX_HIDDEN_LAYERS) return WGG_EINVAL;
  e->n = c->n_enc_hidden;
  e->Z = c->latent_dim;
  e->dims[0] = c->seq_length * c->input_dim;
  int64_t off = 0, aw = 0;
  for (int i = 0; i < e->n; ++i) {
    e->dims[i + 1] = c->enc_hidden_dims[i];
    e->off_w[i] = off; off += (int64_t)e->dims[i + 1] * e->dims[i];
    e->off_b[i] = off; off += e->dims[i + 1];
    e->act_off[i] = aw; aw += e->dims[i + 1];
  }
  const int last = e->dims[e->n];
  e->off_wmu = off; off += (int64_t)e->Z * last;
  e->off_bmu = off; off += e->Z;
  e->off_wlv = off; off += (int64_t)e->Z * last;
  e->off_blv = off; off += e->Z;
  e->total = off;
  e->act_width = aw;
  return WGG_OK;
}

__global__ void reparam_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                               float* __restrict__ z, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    z[i] = mu[i] + eps[i] * expf(0.5f * lv[i]);
}

// Fused latent head: fc_mu and fc_log_var (models.py:70-71), the reparameterisation z = mu + eps * exp(0.5 log_var)
// (:84-86) and the partial sums of the KL term  -0.5 * sum_j (1 + lv - mu^2 - exp(lv))  (losses.py:174-175) in ONE
// launch: both weight matrices (2 * Z * K floats, 8 KB for the default model) are staged in shared memory, transposed to
// [k][j] so that consecutive threads read consecutive words; a block walks over groups of R rows of the last hidden
// activation (one warp-coalesced load per row), thread = (row, latent j).  The KL partials are block sums over a fixed
// partition, summed in fixed order by finalize (deterministic).  kl_partial may be null (value not wanted).
constexpr int kHeadRows = 8;
__global__ void __launch_bounds__(256) enc_head_fused_kernel(const float* __restrict__ h, const float* __restrict__ wmu,
                                                             const float* __restrict__ bmu, const float* __restrict__ wlv,
                                                             const float* __restrict__ blv, const float* __restrict__ eps,
                                                             int64_t B, int K, int Z, float* __restrict__ mu,
                                                             float* __restrict__ lv, float* __restrict__ z,
                                                             float* __restrict__ kl_partial) {
  extern __shared__ float sm[];
  float* s_wmu = sm;                   // [K][Z]
  float* s_wlv = s_wmu + (size_t)K * Z;
  float* s_h = s_wlv + (size_t)K * Z;  // [kHeadRows][K]
  __shared__ float red[33];
  for (int i = threadIdx.x; i < K * Z; i += blockDim.x) {
    const int j = i / K, k = i - j * K;  // global [j][k] -> shared [k][j]
    s_wmu[k * Z + j] = __ldg(wmu + i);
    s_wlv[k * Z + j] = __ldg(wlv + i);
  }
  float acc = 0.f;
  const int64_t groups = (B + kHeadRows - 1) / kHeadRows;
  for (int64_t gidx = blockIdx.x; gidx < groups; gidx += gridDim.x) {
    const int64_t r0 = gidx * kHeadRows;
    __syncthreads();  // weights staged / previous group's rows consumed
    for (int i = threadIdx.x; i < kHeadRows * K; i += blockDim.x) {
      const int64_t r = r0 + i / K;
      s_h[i] = r < B ? __ldg(h + r * K + (i % K)) : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kHeadRows * Z; e += blockDim.x) {
      const int rr = e / Z, j = e - rr * Z;
      const int64_t r = r0 + rr;
      if (r >= B) continue;
      const float* hr = s_h + rr * K;
      float m = __ldg(bmu + j), l = __ldg(blv + j);
      for (int k = 0; k < K; ++k) {
        const float hv = hr[k];
        m = fmaf(hv, s_wmu[k * Z + j], m);
        l = fmaf(hv, s_wlv[k * Z + j], l);
      }
      const int64_t o = r * Z + j;
      const float el = expf(l);
      mu[o] = m;
      lv[o] = l;
      z[o] = m + __ldg(eps + o) * expf(0.5f * l);
      acc += 1.f + l - m * m - el;
    }
  }
  const float sacc = block_sum(acc, red);
  if (kl_partial && threadIdx.x == 0) kl_partial[blockIdx.x] = -0.5f * sacc;
}

// dmu_t = dz + dmu + gk * mu ; dlv_t = dz * eps * 0.5 * exp(0.5 lv) + dlv + gk * 0.5 * (exp(lv) - 1)
// with gk = dkl * kl_coef the upstream gradient of the fused KL output (null: no KL term)
__global__ void enc_head_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ dmu,
                                    const float* __restrict__ dlv, const float* __restrict__ eps,
                                    const float* __restrict__ lv, const float* __restrict__ mu,
                                    const float* __restrict__ dkl, float kl_coef, float* __restrict__ dmu_t,
                                    float* __restrict__ dlv_t, int64_t n) {
  const float gk = dkl ? __ldg(dkl) * kl_coef : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = dz ? dz[i] : 0.f;
    float a = g + (dmu ? dmu[i] : 0.f);
    float b = g * eps[i] * 0.5f * expf(0.5f * lv[i]) + (dlv ? dlv[i] : 0.f);
    if (dkl) {
      a = fmaf(gk, mu[i], a);
      b = fmaf(gk * 0.5f, expf(lv[i]) - 1.f, b);
    }
    dmu_t[i] = a;
    dlv_t[i] = b;
  }
}

int linear_fwd(wgg_ctx* ctx, const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc,
               int64_t M, int N, int K, int act, cudaStream_t st) {
  GemmP p;
  p.A = A; p.M = M; p.K = K; p.sam = lda; p.sak = 1;
  p.B = W; p.N = N; p.sbk = 1; p.sbn = K;
  p.C = C; p.scm = ldc; p.scn = 1;
  p.bias = bias; p.act = act; p.force_fp32 = 1; p.tag = "gemm_kernel/linear_fwd";
  return gemm_launch(ctx, p, st);
}

}  // namespace

// shared with disc.cu
int wgg_linear_fwd(wgg_ctx* ctx, const float* A, int64_t lda, const float* W, const float* bias, float* C,
                   int64_t ldc, int64_t M, int N, int K, int act, cudaStream_t st) {
  return linear_fwd(ctx, A, lda, W, bias, C, ldc, M, N, K, act, st);
}

// dW (N x K) (+)= dY^T (M x N)^T * A (M x K), deterministic split over M
int wgg_linear_wgrad(wgg_ctx* ctx, const float* dY, int64_t ldy, const float* A, int64_t lda, float* dW, float* db, int64_t M,
                     int N, int K, int accumulate, float* part, cudaStream_t st) {
  // db (optional): bias gradient += column sums of dY, carried by the same GEMM (row sums of its A operand dY^T)
  GemmP p;
  p.rowsum = db;
  p.A = dY; p.M = N; p.K = M; p.sam = 1; p.sak = ldy;
  p.B = A; p.N = K; p.sbk = lda; p.sbn = 1;
  p.C = dW; p.scm = K; p.scn = 1; p.accumulate = accumulate;
  p.splitk = gemm_choose_splitk(ctx, p.M, p.N, p.K, 1);
  p.partial = part; p.force_fp32 = 1; p.tag = "gemm_kernel/linear_wgrad";
  return gemm_launch(ctx, p, st);
}

// dA (M x K) (+)= dY (M x N) * W (N x K)
int wgg_linear_dgrad(wgg_ctx* ctx, const float* dY, int64_t ldy, const float* W, float* dA, int64_t lda, int64_t M,
                     int N, int K, int accumulate, cudaStream_t st) {
  GemmP p;
  p.A = dY; p.M = M; p.K = N; p.sam = ldy; p.sak = 1;
  p.B = W; p.N = K; p.sbk = K; p.sbn = 1;
  p.C = dA; p.scm = lda; p.scn = 1; p.accumulate = accumulate; p.force_fp32 = 1; p.tag = "gemm_kernel/linear_dgrad";
  return gemm_launch(ctx, p, st);
}

extern "C" int64_t wgg_encoder_param_floats(const wgg_model_cfg* cfg) {
  EncLayout e;
  return enc_layout(cfg, &e) == WGG_OK ? e.total : -1;
}

extern "C" int64_t wgg_encoder_stash_floats(const wgg_model_cfg* cfg, int64_t B) {
  EncLayout e;
  return enc_layout(cfg, &e) == WGG_OK ? e.act_width * B : -1;
}

static int64_t enc_max_wn(const EncLayout& e) {
  int64_t m = 0;
  for (int i = 0; i < e.n; ++i) {
    const int64_t v = (int64_t)e.dims[i + 1] * e.dims[i];
    if (v > m) m = v;
  }
  return m;
}

extern "C" int64_t wgg_encoder_workspace_floats(const wgg_model_cfg* cfg, int64_t B) {
  EncLayout e;
  if (enc_layout(cfg, &e) != WGG_OK) return -1;
  int maxd = e.Z;
  for (int i = 1; i <= e.n; ++i) maxd = e.dims[i] > maxd ? e.dims[i] : maxd;
  return 2 * B * e.Z + 2 * B * maxd + gemm_splitk_ws_floats(1, enc_max_wn(e), 1) + colsum_ws_floats(maxd, 1);
}

extern "C" int wgg_encoder_forward_kl(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* x,
                                      const float* eps, int64_t B, float* z, float* mu, float* log_var, float* stash,
                                      float* kl, void* stream) {
  EncLayout e;
  if (!ctx) return WGG_EINVAL;
  if (enc_layout(cfg, &e) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "encoder: bad config%s");
  if (!stash) return wgg_fail(ctx, WGG_EINVAL, "encoder_forward: stash (activation buffer) is required%s");
  if (B <= 0) return WGG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float* in = x;
  int64_t ldin = e.dims[0];
  for (int i = 0; i < e.n; ++i) {
    float* out = stash + e.act_off[i] * B;
    WGG_TRY(linear_fwd(ctx, in, ldin, params + e.off_w[i], params + e.off_b[i], out, e.dims[i + 1], B, e.dims[i + 1],
                       e.dims[i], ACT_LEAKY, st));
    in = out;
    ldin = e.dims[i + 1];
  }
  const int last = e.dims[e.n];
  const size_t hsmem = ((size_t)2 * e.Z * last + (size_t)kHeadRows * last) * sizeof(float);
  if (hsmem <= 48 * 1024) {
    float* partial = kl ? wgg_next_partial(ctx) : nullptr;
    const int64_t groups = cdiv64(B, kHeadRows);
    const int nb = (int)(groups < kRedBlocks ? groups : kRedBlocks);
    enc_head_fused_kernel<<<nb, 256, hsmem, st>>>(in, params + e.off_wmu, params + e.off_bmu, params + e.off_wlv,
                                                  params + e.off_blv, eps, B, last, e.Z, mu, log_var, z, partial);
    WGG_CHECK_LAUNCH(ctx, "enc_head_fused_kernel");
    if (kl) WGG_TRY(wgg_loss_finalize(ctx, partial, nb, (float)(1.0 / (double)B), 0, kl, st));
    return WGG_OK;
  }
  // latent heads too large for the shared-memory kernel: two GEMMs, the elementwise reparameterisation, the KL reduction
  WGG_TRY(linear_fwd(ctx, in, last, params + e.off_wmu, params + e.off_bmu, mu, e.Z, B, e.Z, last, ACT_NONE, st));
  WGG_TRY(linear_fwd(ctx, in, last, params + e.off_wlv, params + e.off_blv, log_var, e.Z, B, e.Z, last, ACT_NONE, st));
  reparam_kernel<<<ew_blocks(B * e.Z), 256, 0, st>>>(mu, log_var, eps, z, B * e.Z);
  WGG_CHECK_LAUNCH(ctx, "reparam_kernel");
  if (kl) return wgg_kl(ctx, mu, log_var, B, e.Z, 1.f, 0, kl, stream);
  return WGG_OK;
}

extern "C" int wgg_encoder_forward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* x,
                                   const float* eps, int64_t B, float* z, float* mu, float* log_var, float* stash,
                                   void* stream) {
  return wgg_encoder_forward_kl(ctx, cfg, params, x, eps, B, z, mu, log_var, stash, nullptr, stream);
}

extern "C" int wgg_encoder_backward_kl(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* x,
                                       const float* eps, const float* mu, const float* log_var, int64_t B,
                                       const float* stash, const float* dz, const float* dmu, const float* dlog_var,
                                       const float* dkl, float* dparams, float* dx, float* ws, int64_t ws_floats,
                                       void* stream) {
  EncLayout e;
  if (!ctx) return WGG_EINVAL;
  if (enc_layout(cfg, &e) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "encoder: bad config%s");
  if (B <= 0) return WGG_OK;
  if (!ws || ws_floats < wgg_encoder_workspace_floats(cfg, B))
    return wgg_fail(ctx, WGG_EWORKSPACE, "encoder_backward: workspace too small%s");
  cudaStream_t st = (cudaStream_t)stream;
  int maxd = e.Z;
  for (int i = 1; i <= e.n; ++i) maxd = e.dims[i] > maxd ? e.dims[i] : maxd;
  float* dmu_t = ws;
  float* dlv_t = dmu_t + B * e.Z;
  float* dh = dlv_t + B * e.Z;
  float* dh2 = dh + B * maxd;
  float* part = dh2 + B * maxd;
  float* csws = part + gemm_splitk_ws_floats(1, enc_max_wn(e), 1);
  if (dkl && !mu) return wgg_fail(ctx, WGG_EINVAL, "encoder_backward: the KL gradient needs mu%s");
  enc_head_bwd_kernel<<<ew_blocks(B * e.Z), 256, 0, st>>>(dz, dmu, dlog_var, eps, log_var, mu, dkl, (float)(1.0 / (double)B), dmu_t,
                                                          dlv_t, B * e.Z);
  WGG_CHECK_LAUNCH(ctx, "enc_head_bwd_kernel");
  const int last = e.dims[e.n];
  const float* hl = stash + e.act_off[e.n - 1] * B;
  WGG_TRY(wgg_linear_wgrad(ctx, dmu_t, e.Z, hl, last, dparams + e.off_wmu, dparams + e.off_bmu, B, e.Z, last, 1, part, st));
  WGG_TRY(wgg_linear_wgrad(ctx, dlv_t, e.Z, hl, last, dparams + e.off_wlv, dparams + e.off_blv, B, e.Z, last, 1, part, st));
  WGG_TRY(wgg_linear_dgrad(ctx, dmu_t, e.Z, params + e.off_wmu, dh, last, B, e.Z, last, 0, st));
  WGG_TRY(wgg_linear_dgrad(ctx, dlv_t, e.Z, params + e.off_wlv, dh, last, B, e.Z, last, 1, st));
  for (int i = e.n - 1; i >= 0; --i) {
    const int N = e.dims[i + 1], K = e.dims[i];
    const float* act = stash + e.act_off[i] * B;
    const float* in = i == 0 ? x : stash + e.act_off[i - 1] * B;
    WGG_TRY(leaky_bwd_launch(ctx, act, dh, nullptr, B * N, st));
    WGG_TRY(wgg_linear_wgrad(ctx, dh, N, in, K, dparams + e.off_w[i], dparams + e.off_b[i], B, N, K, 1, part, st));
    if (i > 0) {
      WGG_TRY(wgg_linear_dgrad(ctx, dh, N, params + e.off_w[i], dh2, K, B, N, K, 0, st));
      float* t = dh; dh = dh2; dh2 = t;
    } else if (dx) {
      WGG_TRY(wgg_linear_dgrad(ctx, dh, N, params + e.off_w[i], dx, K, B, N, K, 0, st));
    }
  }
  return WGG_OK;
}

extern "C" int wgg_encoder_backward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* x,
                                    const float* eps, const float* log_var, int64_t B, const float* stash,
                                    const float* dz, const float* dmu, const float* dlog_var, float* dparams,
                                    float* dx, float* ws, int64_t ws_floats, void* stream) {
  return wgg_encoder_backward_kl(ctx, cfg, params, x, eps, nullptr, log_var, B, stash, dz, dmu, dlog_var, nullptr, dparams, dx,
                                 ws, ws_floats, stream);
}

extern "C" int wgg_linear(wgg_ctx* ctx, const float* A, const float* W, const float* bias, float* C, int64_t M,
                          int32_t N, int32_t K, int act, void* stream) {
  if (!ctx) return WGG_EINVAL;
  return linear_fwd(ctx, A, K, W, bias, C, N, M, N, K, act, (cudaStream_t)stream);
}

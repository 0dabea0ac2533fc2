// Spectral-normalised discriminators (Conv1D "TemporalDiscriminator" - the default - and the MLP one),
// forward and backward, including the spectral_norm power iteration and its gradient.
// Replaces TemporalDiscriminator.forward/get_all_features (src/gan/models.py:293-353),
// Discriminator.forward/get_all_features (:202-243) and torch.nn.utils.spectral_norm's pre-forward
// hook (torch/nn/utils/spectral_norm.py:62-114) for every layer they wrap (models.py:194,198,270-291).
//
// Layout: activations are channel-last (B,T,C) so a conv1d is an implicit GEMM over a sliding window
// of contiguous floats (see GemmP::conv_mode).  The stash of one forward call holds every post-LeakyReLU
// activation (= the feature-matching features) plus the pooled tensor: [B x width] blocks back to back.
// `sn` holds, per layer, W_orig/sigma in kernel-ready layouts plus the (u, v, sigma) the call used.
#include "common.cuh"

namespace {

constexpr int kMaxLayers = WGG_MAX_HIDDEN_LAYERS + 2;
constexpr int kPoolBins = 8;  // nn.AdaptiveAvgPool1d(8), models.py:278

struct DLayer {
  int rows, cols;       // weight matrix as spectral_norm sees it: (out, in*k)
  int is_conv, Cin, ks, pad;
  int64_t off_b, off_w;  // into flat params
  int64_t off_u, off_v;  // into flat uv
  int64_t sn_wf, sn_wb, sn_u, sn_v, sn_sigma;  // into sn
  int64_t sn_tcf, sn_tcd;                      // tcgen05 weight images (conv layers): forward / backward-data
  int tc_kf, tc_nd;                            // forward image K chunks; backward-data image N (>= 16)
  int64_t g_off;         // into the effective-weight-gradient scratch
  int in_width, out_width;  // per-sample activation widths
  int64_t in_off, out_off;  // per-sample offsets into stash (in_off < 0: network input x)
  int feat;                 // feature index of the output activation or -1
};

struct DLayout {
  int nl, temporal, T, C;
  DLayer L[kMaxLayers];
  int64_t param_total, uv_total, sn_total, g_total;
  int64_t stash_width;      // per sample
  int64_t pooled_off;       // per-sample offset of pooled block (temporal only)
  int64_t x4_off;           // per-sample offset of the channel-padded input copy (tcgen05 path)
  int first_linear;         // index of the first Linear after the pool (temporal) or -1
  int nfeat;
  int64_t feat_off[kMaxLayers];
  int feat_width[kMaxLayers];
  int max_width;            // widest per-sample activation (incl. input)
  int64_t max_wn;           // largest rows*cols
  int max_rows;
};

int disc_layout(const wgg_model_cfg* c, DLayout* d) {
  if (!c) return WGG_EINVAL;
  memset(d, 0, sizeof(*d));
  d->T = c->seq_length; d->C = c->input_dim; d->temporal = c->use_temporal_disc ? 1 : 0;
  d->first_linear = -1;
  int nl = 0;
  auto add = [&](int rows, int cols, int is_conv, int Cin, int ks) {
    DLayer& l = d->L[nl++];
    l.rows = rows; l.cols = cols; l.is_conv = is_conv; l.Cin = Cin; l.ks = ks; l.pad = (ks - 1) / 2;
  };
  if (d->temporal) {
    add(64, d->C * 5, 1, d->C, 5);   // models.py:270
    add(64, 64 * 5, 1, 64, 5);       // :273
    add(32, 64 * 3, 1, 64, 3);       // :276
    add(128, 32 * kPoolBins, 0, 0, 1);  // :285
    add(64, 128, 0, 0, 1);           // :287
    add(1, 64, 0, 0, 1);             // :290
    d->first_linear = 3;
  } else {
    if (c->n_disc_hidden < 1 || c->n_disc_hidden > WGG_MAX_HIDDEN_LAYERS) return WGG_EINVAL;
    int in = d->T * d->C;
    for (int i = 0; i < c->n_disc_hidden; ++i) { add(c->disc_hidden_dims[i], in, 0, 0, 1); in = c->disc_hidden_dims[i]; }
    add(1, in, 0, 0, 1);             // models.py:198
  }
  d->nl = nl;
  int64_t po = 0, uo = 0, so = 0, go = 0, sw = 0;
  int nf = 0;
  d->max_width = d->T * d->C;
  for (int i = 0; i < nl; ++i) {
    DLayer& l = d->L[i];
    const int64_t wn = (int64_t)l.rows * l.cols;
    l.off_b = po; po += l.rows;       // named_parameters order: bias, weight_orig
    l.off_w = po; po += wn;
    l.off_u = uo; uo += l.rows;
    l.off_v = uo; uo += l.cols;
    l.sn_wf = so; so += wn;
    l.sn_wb = so; so += l.is_conv ? wn : 0;
    l.sn_u = so; so += l.rows;
    l.sn_v = so; so += l.cols;
    l.sn_sigma = so; so += 4;         // keep 16-byte alignment of the following blocks
    so = (so + 3) & ~(int64_t)3;
    l.sn_tcf = l.sn_tcd = so; l.tc_kf = 0; l.tc_nd = 0;
    if (l.is_conv) {
      const int CinC = (l.Cin + 3) / 4;
      const int taps_p = CinC == 1 ? l.ks + (l.ks & 1) : l.ks;
      l.tc_kf = CinC == 1 ? taps_p : l.ks * CinC;
      l.tc_nd = CinC * 4 < 16 ? 16 : CinC * 4;
      l.sn_tcf = so; so += (int64_t)l.tc_kf * 4 * l.rows;
      l.sn_tcd = so; so += (int64_t)l.ks * l.rows * l.tc_nd;
    }
    l.g_off = go; go += wn;
    if (wn > d->max_wn) d->max_wn = wn;
    if (l.rows > d->max_rows) d->max_rows = l.rows;
    // activation bookkeeping
    if (i == 0) { l.in_off = -1; l.in_width = d->T * d->C; }
    else if (i == d->first_linear) { l.in_off = d->pooled_off; l.in_width = l.cols; }
    else { l.in_off = d->L[i - 1].out_off; l.in_width = d->L[i - 1].out_width; }
    l.out_width = l.is_conv ? d->T * l.rows : l.rows;
    l.feat = -1;
    if (i < nl - 1) {
      l.out_off = sw; sw += l.out_width;
      l.feat = nf;
      d->feat_off[nf] = l.out_off; d->feat_width[nf] = l.out_width; ++nf;
      if (l.out_width > d->max_width) d->max_width = l.out_width;
      if (d->temporal && i + 1 == d->first_linear) {
        d->pooled_off = sw; sw += l.rows * kPoolBins;
      }
    } else {
      l.out_off = -1;
    }
  }
  d->x4_off = sw; sw += (int64_t)d->T * 4;
  d->nfeat = nf;
  d->param_total = po; d->uv_total = uo; d->sn_total = so; d->g_total = go; d->stash_width = sw;
  return WGG_OK;
}

// ---------------------------------------------------------------------------------------------
// spectral norm: one CTA per layer (weights are <= 74k floats).  torch/nn/utils/spectral_norm.py:92-114
// ---------------------------------------------------------------------------------------------
struct SnArgs {
  int nl;
  int rows[kMaxLayers], cols[kMaxLayers], is_conv[kMaxLayers], Cin[kMaxLayers], ks[kMaxLayers];
  int64_t off_w[kMaxLayers], off_u[kMaxLayers], off_v[kMaxLayers];
  int64_t sn_wf[kMaxLayers], sn_wb[kMaxLayers], sn_u[kMaxLayers], sn_v[kMaxLayers], sn_sigma[kMaxLayers];
  int64_t g_off[kMaxLayers];
  int64_t sn_tcf[kMaxLayers], sn_tcd[kMaxLayers];
  int tc_kf[kMaxLayers], tc_nd[kMaxLayers];
};

__global__ void __launch_bounds__(1024) sn_kernel(SnArgs a, const float* __restrict__ params, float* __restrict__ uv,
                                                  float* __restrict__ sn, int training, int with_output) {
  extern __shared__ float sm[];
  const int l = blockIdx.x;
  if (l == a.nl - 1 && !with_output) return;  // get_all_features never calls output_layer (models.py:319-353)
  const int rows = a.rows[l], cols = a.cols[l];
  const int NT = blockDim.x;
  float* u_s = sm;
  float* v_s = u_s + rows;
  float* wv_s = v_s + cols;
  float* red = wv_s + rows;        // 33 floats
  float* vpart = red + 40;         // [4][cols] partial sums of W^T u over row slices
  const float* __restrict__ W = params + a.off_w[l];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
  for (int r = tid; r < rows; r += NT) u_s[r] = uv[a.off_u[l] + r];
  for (int c = tid; c < cols; c += NT) v_s[c] = uv[a.off_v[l] + c];
  __syncthreads();
  if (training) {
    // v = normalize(W^T u): 4 row slices x column threads, combined in fixed order (deterministic)
    const int slice = tid / (NT / 4), ct = tid % (NT / 4);
    const int r0 = (rows * slice) / 4, r1 = (rows * (slice + 1)) / 4;
    for (int c = ct; c < cols; c += NT / 4) {
      float s = 0.f;
#pragma unroll 8
      for (int r = r0; r < r1; ++r) s = fmaf(__ldg(W + (int64_t)r * cols + c), u_s[r], s);
      vpart[slice * cols + c] = s;
    }
    __syncthreads();
    float ss = 0.f;
    for (int c = tid; c < cols; c += NT) {
      const float s = (vpart[c] + vpart[cols + c]) + (vpart[2 * cols + c] + vpart[3 * cols + c]);
      v_s[c] = s;
      ss += s * s;
    }
    const float nv = sqrtf(block_sum(ss, red));
    const float dv = fmaxf(nv, 1e-12f);
    for (int c = tid; c < cols; c += NT) v_s[c] = v_s[c] / dv;
    __syncthreads();
  }
  // wv = W v   (warp per row)
  for (int r = warp; r < rows; r += nwarp) {
    float s = 0.f;
#pragma unroll 4
    for (int c = lane; c < cols; c += 32) s = fmaf(__ldg(W + (int64_t)r * cols + c), v_s[c], s);
    s = warp_sum(s);
    if (lane == 0) wv_s[r] = s;
  }
  __syncthreads();
  if (training) {
    float ss = 0.f;
    for (int r = tid; r < rows; r += NT) ss += wv_s[r] * wv_s[r];
    const float nu = sqrtf(block_sum(ss, red));
    const float du = fmaxf(nu, 1e-12f);
    for (int r = tid; r < rows; r += NT) u_s[r] = wv_s[r] / du;
    __syncthreads();
    for (int r = tid; r < rows; r += NT) uv[a.off_u[l] + r] = u_s[r];
    for (int c = tid; c < cols; c += NT) uv[a.off_v[l] + c] = v_s[c];
  }
  float sp = 0.f;
  for (int r = tid; r < rows; r += NT) sp += u_s[r] * wv_s[r];
  const float sigma = block_sum(sp, red);
  for (int r = tid; r < rows; r += NT) sn[a.sn_u[l] + r] = u_s[r];
  for (int c = tid; c < cols; c += NT) sn[a.sn_v[l] + c] = v_s[c];
  if (tid == 0) sn[a.sn_sigma[l]] = sigma;
}

// W / sigma in the layouts the compute kernels read (second launch: wide, so that the single-CTA power iteration
// above stays short).  grid (layers, kSnImageSplit)
constexpr int kSnImageSplit = 16;
__global__ void __launch_bounds__(256) sn_images_kernel(SnArgs a, const float* __restrict__ params, float* __restrict__ sn,
                                                        int with_output) {
  const int l = blockIdx.x;
  if (l == a.nl - 1 && !with_output) return;
  const int rows = a.rows[l], cols = a.cols[l];
  const float* __restrict__ W = params + a.off_w[l];
  const int tid = blockIdx.y * blockDim.x + threadIdx.x, NT = blockDim.x * gridDim.y;
  const float sigma = sn[a.sn_sigma[l]];
  float* wf = sn + a.sn_wf[l];
  float* wb = sn + a.sn_wb[l];
  const int n = rows * cols;
  if (a.is_conv[l]) {
    const int ks = a.ks[l], Cin = a.Cin[l];
    for (int idx = tid; idx < n; idx += NT) {
      const int r = idx / cols, c = idx % cols;
      const int ci = c / ks, k = c % ks;
      const float w = W[idx] / sigma;
      wf[r * cols + k * Cin + ci] = w;                       // [co][(k,ci)]   forward / weight-grad layout
      wb[((ks - 1 - k) * rows + r) * Cin + ci] = w;          // [(k',co)][ci]  backward-data layout (flipped taps)
    }
    // tcgen05 images (TF32, UMMA K-major core-matrix order [K/4][N/8][8][4]); see conv_tc.cu
    const int CinC = (Cin + 3) / 4, Cin4 = CinC * 4;
    const int kf = a.tc_kf[l] * 4;                      // forward K (tap-major, channels padded to Cin4)
    float* tf = sn + a.sn_tcf[l];
    for (int idx = tid; idx < rows * kf; idx += NT) {
      const int co = idx / kf, kk = idx % kf;
      const int tap = kk / Cin4, ci = kk % Cin4;
      const float w = (tap < ks && ci < Cin) ? W[(co * Cin + ci) * ks + tap] / sigma : 0.f;
      uint32_t t32;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t32) : "f"(w));
      tf[(kk >> 2) * (rows / 8) * 32 + (co >> 3) * 32 + (co & 7) * 4 + (kk & 3)] = __uint_as_float(t32);
    }
    const int nd = a.tc_nd[l], kd = ks * rows;          // backward-data: N = input channels, K = (flipped tap, co)
    float* td = sn + a.sn_tcd[l];
    for (int idx = tid; idx < nd * kd; idx += NT) {
      const int ci = idx / kd, kk = idx % kd;
      const int tp = kk / rows, co = kk % rows;
      const float w = ci < Cin ? W[(co * Cin + ci) * ks + (ks - 1 - tp)] / sigma : 0.f;
      uint32_t t32;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t32) : "f"(w));
      td[(kk >> 2) * (nd / 8) * 32 + (ci >> 3) * 32 + (ci & 7) * 4 + (kk & 3)] = __uint_as_float(t32);
    }
  } else {
    for (int idx = tid; idx < n; idx += NT) wf[idx] = W[idx] / sigma;
  }
}

// dW_orig += G/sigma - (<G, W_orig>/sigma^2) u v^T, G given in the forward layout of that layer.
// Two wide launches, grid (layers, kSnGradSplit): partial inner products <G, W_orig>, then the update (the partials
// are combined in fixed order by every block, so the result is deterministic).
constexpr int kSnGradSplit = 16;
__device__ __forceinline__ int sn_gidx(int idx, int cols, int ks, int Cin, int conv) {
  if (!conv) return idx;
  const int r = idx / cols, c = idx % cols;
  return r * cols + (c % ks) * Cin + c / ks;
}
__global__ void __launch_bounds__(256) sn_inner_kernel(SnArgs a, const float* __restrict__ params,
                                                       const float* __restrict__ G, float* __restrict__ partial,
                                                       int with_output) {
  __shared__ float red[33];
  const int l = blockIdx.x;
  if (l == a.nl - 1 && !with_output) return;  // features-only call: the output layer never ran
  const int rows = a.rows[l], cols = a.cols[l], n = rows * cols;
  const float* __restrict__ W = params + a.off_w[l];
  const float* __restrict__ g = G + a.g_off[l];
  const int ks = a.ks[l], Cin = a.Cin[l], conv = a.is_conv[l];
  float ip = 0.f;
#pragma unroll 4
  for (int idx = blockIdx.y * 256 + threadIdx.x; idx < n; idx += 256 * kSnGradSplit)
    ip = fmaf(__ldg(g + sn_gidx(idx, cols, ks, Cin, conv)), __ldg(W + idx), ip);
  const float inner = block_sum(ip, red);
  if (threadIdx.x == 0) partial[l * kSnGradSplit + blockIdx.y] = inner;
}
__global__ void __launch_bounds__(256) sn_grad_kernel(SnArgs a, const float* __restrict__ sn, const float* __restrict__ G,
                                                      const float* __restrict__ partial, float* __restrict__ dparams,
                                                      int with_output) {
  const int l = blockIdx.x;
  if (l == a.nl - 1 && !with_output) return;
  const int rows = a.rows[l], cols = a.cols[l], n = rows * cols;
  const float* __restrict__ g = G + a.g_off[l];
  const int ks = a.ks[l], Cin = a.Cin[l], conv = a.is_conv[l];
  float inner = 0.f;
#pragma unroll
  for (int i = 0; i < kSnGradSplit; ++i) inner += partial[l * kSnGradSplit + i];
  const float sigma = sn[a.sn_sigma[l]];
  const float k2 = inner / (sigma * sigma);
  const float* u = sn + a.sn_u[l];
  const float* v = sn + a.sn_v[l];
  float* dW = dparams + a.off_w[l];
  for (int idx = blockIdx.y * 256 + threadIdx.x; idx < n; idx += 256 * kSnGradSplit) {
    const int r = idx / cols, c = idx % cols;
    dW[idx] += g[sn_gidx(idx, cols, ks, Cin, conv)] / sigma - k2 * u[r] * v[c];
  }
}

// pooled[b][c*8+bin] = mean_{t in bin} a[b][t][c]          (AdaptiveAvgPool1d(8) + flatten, models.py:312-315)
__global__ void pool_fwd_kernel(const float* __restrict__ a, float* __restrict__ pooled, int64_t B, int T, int C) {
  const int64_t n = B * C * kPoolBins;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int bin = (int)((i / C) % kPoolBins);
    const int64_t b = i / ((int64_t)C * kPoolBins);
    const int s = (bin * T) / kPoolBins;
    const int e = ((bin + 1) * T + kPoolBins - 1) / kPoolBins;
    float acc = 0.f;
    for (int t = s; t < e; ++t) acc += __ldg(a + (b * T + t) * C + c);
    pooled[b * C * kPoolBins + c * kPoolBins + bin] = acc / (float)(e - s);
  }
}

__global__ void pool_bwd_kernel(const float* __restrict__ dpool, float* __restrict__ da, int64_t B, int T, int C) {
  const int64_t n = B * T * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int t = (int)((i / C) % T);
    const int64_t b = i / ((int64_t)C * T);
    float acc = 0.f;
#pragma unroll
    for (int bin = 0; bin < kPoolBins; ++bin) {
      const int s = (bin * T) / kPoolBins;
      const int e = ((bin + 1) * T + kPoolBins - 1) / kPoolBins;
      if (t >= s && t < e) acc += __ldg(dpool + b * C * kPoolBins + c * kPoolBins + bin) / (float)(e - s);
    }
    da[i] = acc;
  }
}

// (B,T,C) -> (B,C,T)
__global__ void transpose_tc_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t B, int T, int C) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    tile[i][tx] = (t < T && c < C) ? in[(b * T + t) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    if (t < T && c < C) out[(b * C + c) * T + t] = tile[tx][i];
  }
}

// tcgen05 conv path: TemporalDiscriminator with T a multiple of 128 (= UMMA M; T = 256 in the scaled regime of
// BASELINE configs[3]) and <= 4 input channels, in tf32 mode
bool disc_use_tc(const wgg_ctx* ctx, const DLayout& d) {
  return ctx->math_mode == 1 && d.temporal && conv_tc_seq_ok(d.T) && d.C <= 4;
}

void fill_sn_args(const DLayout& d, SnArgs* a) {
  a->nl = d.nl;
  for (int i = 0; i < d.nl; ++i) {
    const DLayer& l = d.L[i];
    a->rows[i] = l.rows; a->cols[i] = l.cols; a->is_conv[i] = l.is_conv; a->Cin[i] = l.Cin; a->ks[i] = l.ks;
    a->off_w[i] = l.off_w; a->off_u[i] = l.off_u; a->off_v[i] = l.off_v;
    a->sn_wf[i] = l.sn_wf; a->sn_wb[i] = l.sn_wb; a->sn_u[i] = l.sn_u; a->sn_v[i] = l.sn_v;
    a->sn_sigma[i] = l.sn_sigma; a->g_off[i] = l.g_off;
    a->sn_tcf[i] = l.sn_tcf; a->sn_tcd[i] = l.sn_tcd; a->tc_kf[i] = l.tc_kf; a->tc_nd[i] = l.tc_nd;
  }
}

}  // namespace

int disc_feature_table(const wgg_model_cfg* cfg, int64_t B, FeatTable* ft) {
  DLayout d;
  if (disc_layout(cfg, &d) != WGG_OK) return WGG_EINVAL;
  ft->n = d.nfeat;
  for (int k = 0; k < d.nfeat; ++k) {
    ft->off[k] = d.feat_off[k] * B;
    ft->width[k] = d.feat_width[k];
    ft->count[k] = (int64_t)d.feat_width[k] * B;
  }
  return WGG_OK;
}

extern "C" int64_t wgg_disc_param_floats(const wgg_model_cfg* cfg) {
  DLayout d;
  return disc_layout(cfg, &d) == WGG_OK ? d.param_total : -1;
}
extern "C" int64_t wgg_disc_uv_floats(const wgg_model_cfg* cfg) {
  DLayout d;
  return disc_layout(cfg, &d) == WGG_OK ? d.uv_total : -1;
}
extern "C" int64_t wgg_disc_sn_floats(const wgg_model_cfg* cfg) {
  DLayout d;
  return disc_layout(cfg, &d) == WGG_OK ? d.sn_total : -1;
}
extern "C" int64_t wgg_disc_stash_floats(const wgg_model_cfg* cfg, int64_t B) {
  DLayout d;
  return disc_layout(cfg, &d) == WGG_OK ? d.stash_width * B : -1;
}
extern "C" int32_t wgg_disc_num_features(const wgg_model_cfg* cfg) {
  DLayout d;
  return disc_layout(cfg, &d) == WGG_OK ? d.nfeat : -1;
}
extern "C" int64_t wgg_disc_feature_offset(const wgg_model_cfg* cfg, int64_t B, int32_t k) {
  DLayout d;
  if (disc_layout(cfg, &d) != WGG_OK || k < 0 || k >= d.nfeat) return -1;
  return d.feat_off[k] * B;
}
extern "C" int32_t wgg_disc_feature_width(const wgg_model_cfg* cfg, int32_t k) {
  DLayout d;
  if (disc_layout(cfg, &d) != WGG_OK || k < 0 || k >= d.nfeat) return -1;
  return d.feat_width[k];
}
extern "C" int64_t wgg_disc_workspace_floats(const wgg_model_cfg* cfg, int64_t B) {
  DLayout d;
  if (disc_layout(cfg, &d) != WGG_OK) return -1;
  int64_t part = gemm_splitk_ws_floats(1, d.max_wn, 1);
  const int64_t tcp = (int64_t)256 * 64 * 336;  // tcgen05 wgrad partials: <= 256 CTAs x 64 rows x (5*64 + 8 -> 336) cols
  if (d.temporal && tcp > part) part = tcp;
  // 2 gradient ping-pong buffers (+ 2 row-major staging buffers for the tcgen05 path's weight gradients)
  return (d.temporal ? 4 : 2) * B * (int64_t)d.max_width + d.g_total + part + colsum_ws_floats(d.max_rows, 1);
}

extern "C" int wgg_disc_spectral(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, float* uv, int training,
                                 int with_output_layer, float* sn, void* stream) {
  DLayout d;
  if (!ctx) return WGG_EINVAL;
  if (disc_layout(cfg, &d) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "disc: bad config%s");
  SnArgs a;
  fill_sn_args(d, &a);
  int maxr = 0, maxc = 0;
  for (int i = 0; i < d.nl; ++i) { maxr = d.L[i].rows > maxr ? d.L[i].rows : maxr; maxc = d.L[i].cols > maxc ? d.L[i].cols : maxc; }
  const size_t smem = (size_t)(2 * maxr + maxc + 80 + 4 * maxc) * sizeof(float);
  if (smem > 48 * 1024) return wgg_fail(ctx, WGG_EUNSUPPORTED, "disc_spectral: layer too wide for the SN kernel%s");
  sn_kernel<<<d.nl, 1024, smem, (cudaStream_t)stream>>>(a, params, uv, sn, training, with_output_layer);
  WGG_CHECK_LAUNCH(ctx, "sn_kernel");
  sn_images_kernel<<<dim3(d.nl, kSnImageSplit), 256, 0, (cudaStream_t)stream>>>(a, params, sn, with_output_layer);
  WGG_CHECK_LAUNCH(ctx, "sn_images_kernel");
  return WGG_OK;
}

extern "C" int wgg_disc_forward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* sn,
                                const float* x, int64_t B, float* score, float* stash, void* stream) {
  DLayout d;
  if (!ctx) return WGG_EINVAL;
  if (disc_layout(cfg, &d) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "disc: bad config%s");
  if (!stash) return wgg_fail(ctx, WGG_EINVAL, "disc_forward: stash is required%s");
  if (B <= 0) return WGG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int last = d.nl - 1;
  const bool tc = disc_use_tc(ctx, d);
  if (tc) {
    // conv stack on tcgen05: activations channel-chunked [B][C/4][T][4] (see conv_tc.cu)
    float* x4 = stash + d.x4_off * B;
    WGG_TRY(pack_x4_launch(ctx, x, x4, B, d.T, d.C, st));
    const float* in = x4;
    for (int i = 0; i < d.first_linear; ++i) {
      const DLayer& l = d.L[i];
      float* out = stash + l.out_off * B;
      WGG_TRY(conv_tc_fwd_launch(ctx, in, sn + l.sn_tcf, params + l.off_b, out, nullptr, nullptr, B, d.T, (l.Cin + 3) / 4, l.ks,
                                 l.pad, l.rows, 0, "conv_tc_fwd_kernel/fwd", st));
      in = out;
    }
    WGG_TRY(pool_fwd_chunk_launch(ctx, in, stash + d.pooled_off * B, B, d.T, d.L[d.first_linear - 1].rows, st));
  }
  for (int i = tc ? d.first_linear : 0; i < d.nl; ++i) {
    const DLayer& l = d.L[i];
    if (i == last && !score) break;
    const float* in = l.in_off < 0 ? x : stash + l.in_off * B;
    float* out = i == last ? score : stash + l.out_off * B;
    if (l.is_conv) {
      GemmP p;
      p.tag = "gemm_kernel/conv_fwd";
      p.A = in - (int64_t)l.pad * l.Cin; p.M = B * d.T; p.K = l.cols; p.sam = l.Cin; p.sak = 1;
      p.conv_mode = 1; p.conv_T = d.T; p.conv_Cin = l.Cin; p.conv_pad = l.pad;
      p.B = sn + l.sn_wf; p.N = l.rows; p.sbk = 1; p.sbn = l.cols;
      p.C = out; p.scm = l.rows; p.scn = 1;
      p.bias = params + l.off_b; p.act = ACT_LEAKY;
      WGG_TRY(gemm_launch(ctx, p, st));
      if (d.temporal && i + 1 == d.first_linear) {
        pool_fwd_kernel<<<ew_blocks(B * l.rows * kPoolBins), 256, 0, st>>>(out, stash + d.pooled_off * B, B, d.T, l.rows);
        WGG_CHECK_LAUNCH(ctx, "pool_fwd_kernel");
      }
    } else {
      WGG_TRY(wgg_linear_fwd(ctx, in, l.cols, sn + l.sn_wf, params + l.off_b, out, l.rows, B, l.rows, l.cols,
                             i == last ? ACT_NONE : ACT_LEAKY, st));
    }
  }
  return WGG_OK;
}

extern "C" int wgg_disc_backward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* sn,
                                 const float* x, int64_t B, const float* stash, const float* dscore,
                                 const float* dfeat, float* dparams, float* dx, float* ws, int64_t ws_floats,
                                 void* stream) {
  DLayout d;
  if (!ctx) return WGG_EINVAL;
  if (disc_layout(cfg, &d) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "disc: bad config%s");
  if (B <= 0) return WGG_OK;
  if (!dscore && !dfeat) return wgg_fail(ctx, WGG_EINVAL, "disc_backward: need dscore and/or dfeat%s");
  if (!ws || ws_floats < wgg_disc_workspace_floats(cfg, B))
    return wgg_fail(ctx, WGG_EWORKSPACE, "disc_backward: workspace too small%s");
  cudaStream_t st = (cudaStream_t)stream;
  float* dcur = ws;
  float* dnext = dcur + B * (int64_t)d.max_width;
  float* rm_a = dnext + B * (int64_t)d.max_width;   // row-major staging (tcgen05 path only)
  float* rm_b = d.temporal ? rm_a + B * (int64_t)d.max_width : rm_a;
  float* G = d.temporal ? rm_b + B * (int64_t)d.max_width : rm_a;
  float* part = G + d.g_total;
  float* csws = part + gemm_splitk_ws_floats(1, d.max_wn, 1);
  const int last = d.nl - 1;
  {
    const DLayer& l = d.L[last];
    const float* in = stash + l.in_off * B;
    if (dscore) {
      if (dparams) {
        WGG_TRY(wgg_linear_wgrad(ctx, dscore, 1, in, l.cols, G + l.g_off, dparams + l.off_b, B, 1, l.cols, 0, part, st));
      }
      WGG_TRY(wgg_linear_dgrad(ctx, dscore, 1, sn + l.sn_wf, dcur, l.cols, B, 1, l.cols, 0, st));
    } else {
      WGG_TRY(fill_launch(ctx, dcur, 0.f, B * (int64_t)l.cols, st));
    }
  }
  for (int i = last - 1; i >= 0; --i) {
    const DLayer& l = d.L[i];
    const float* y = stash + l.out_off * B;
    const float* in = l.in_off < 0 ? x : stash + l.in_off * B;
    const float* add = dfeat ? dfeat + l.out_off * B : nullptr;
    WGG_TRY(leaky_bwd_launch(ctx, y, dcur, add, B * (int64_t)l.out_width, st));
    const bool need_dgrad = (i > 0) || dx;
    if (l.is_conv) {
      const int64_t R = B * d.T;
      if (dparams) {
        GemmP p;  // G[co][(k,ci)] = sum_r dpre[r][co] * window(in)(r,(k,ci))
        p.tag = "gemm_kernel/conv_wgrad";
        p.A = dcur; p.M = l.rows; p.K = R; p.sam = 1; p.sak = l.rows;
        p.B = in - (int64_t)l.pad * l.Cin; p.N = l.cols; p.sbk = l.Cin; p.sbn = 1;
        p.conv_mode = 2; p.conv_T = d.T; p.conv_Cin = l.Cin; p.conv_pad = l.pad;
        p.C = G + l.g_off; p.scm = l.cols; p.scn = 1;
        p.splitk = gemm_choose_splitk(ctx, p.M, p.N, p.K, 1); p.partial = part;
        WGG_TRY(gemm_launch(ctx, p, st));
        WGG_TRY(colsum_launch(ctx, dcur, R, l.rows, l.rows, 1, 0, dparams + l.off_b, nullptr, 0, 1, csws, st));
      }
      if (need_dgrad) {
        float* dst = (i == 0) ? dx : dnext;
        const int padb = l.ks - 1 - l.pad;
        GemmP p;  // d_in[r][ci] = sum_{k',co} window(dpre)(r,(k',co)) * Wb[(k',co)][ci]
        p.tag = "gemm_kernel/conv_dgrad";
        p.A = dcur - (int64_t)padb * l.rows; p.M = R; p.K = (int64_t)l.ks * l.rows; p.sam = l.rows; p.sak = 1;
        p.conv_mode = 1; p.conv_T = d.T; p.conv_Cin = l.rows; p.conv_pad = padb;
        p.B = sn + l.sn_wb; p.N = l.Cin; p.sbk = l.Cin; p.sbn = 1;
        p.C = dst; p.scm = l.Cin; p.scn = 1;
        WGG_TRY(gemm_launch(ctx, p, st));
      }
    } else {
      if (dparams) {
        WGG_TRY(wgg_linear_wgrad(ctx, dcur, l.rows, in, l.cols, G + l.g_off, dparams + l.off_b, B, l.rows, l.cols, 0, part, st));
      }
      if (need_dgrad) {
        float* dst = (i == 0) ? dx : dnext;
        WGG_TRY(wgg_linear_dgrad(ctx, dcur, l.rows, sn + l.sn_wf, dst, l.cols, B, l.rows, l.cols, 0, st));
        if (d.temporal && i == d.first_linear && disc_use_tc(ctx, d)) {
          // ---- conv stack backward on tcgen05 (conv_tc.cu): dnext holds d(pooled) ----
          const float* x4 = stash + d.x4_off * B;
          float* dp_hi = dcur;    // d(pre-activation) of the current conv layer, chunk layout
          float* dp_lo = dnext;
          {
            const DLayer& c3 = d.L[i - 1];
            // dnext (d pooled) is consumed into dcur (dpre of the last conv layer)
            WGG_TRY(unpool_leaky_chunk_launch(ctx, dnext, stash + c3.out_off * B, dfeat ? dfeat + c3.out_off * B : nullptr,
                                              dp_hi, B, d.T, c3.rows, st));
          }
          for (int j = i - 1; j >= 0; --j) {
            const DLayer& c = d.L[j];
            const float* cin = j == 0 ? x4 : stash + d.L[j - 1].out_off * B;
            if (dparams)
              WGG_TRY(conv_tc_wgrad_launch(ctx, dp_hi, cin, B, d.T, c.rows, c.Cin, c.ks, c.pad, G + c.g_off, dparams + c.off_b,
                                           part, st));
            if (j > 0) {
              const DLayer& lo = d.L[j - 1];
              WGG_TRY(conv_tc_fwd_launch(ctx, dp_hi, sn + c.sn_tcd, nullptr, dp_lo, stash + lo.out_off * B,
                                         dfeat ? dfeat + lo.out_off * B : nullptr, B, d.T, c.rows / 4, c.ks, c.ks - 1 - c.pad,
                                         c.tc_nd, 1, "conv_tc_fwd_kernel/dgrad", st));
              float* t = dp_hi; dp_hi = dp_lo; dp_lo = t;
            } else if (dx) {
              WGG_TRY(conv_tc_fwd_launch(ctx, dp_hi, sn + c.sn_tcd, nullptr, dx, nullptr, nullptr, B, d.T, c.rows / 4, c.ks,
                                         c.ks - 1 - c.pad, c.tc_nd, 2, "conv_tc_fwd_kernel/dx", st));
            }
          }
          break;  // the conv layers are done
        }
        if (d.temporal && i == d.first_linear) {
          // dnext holds d(pooled); un-pool into the conv activation gradient
          const DLayer& c = d.L[i - 1];
          pool_bwd_kernel<<<ew_blocks(B * d.T * c.rows), 256, 0, st>>>(dnext, dcur, B, d.T, c.rows);
          WGG_CHECK_LAUNCH(ctx, "pool_bwd_kernel");
          continue;  // dcur already holds the gradient for layer i-1's output
        }
      }
    }
    if (i > 0) { float* t = dcur; dcur = dnext; dnext = t; }
  }
  if (dparams) {
    SnArgs a;
    fill_sn_args(d, &a);
    float* ipart = wgg_next_partial(ctx);
    sn_inner_kernel<<<dim3(d.nl, kSnGradSplit), 256, 0, st>>>(a, params, G, ipart, dscore ? 1 : 0);
    WGG_CHECK_LAUNCH(ctx, "sn_inner_kernel");
    sn_grad_kernel<<<dim3(d.nl, kSnGradSplit), 256, 0, st>>>(a, sn, G, ipart, dparams, dscore ? 1 : 0);
    WGG_CHECK_LAUNCH(ctx, "sn_grad_kernel");
  }
  return WGG_OK;
}

namespace {
// stash block of a conv feature <-> the public (B, C*T) layout of get_all_features (h.view(B,-1) of (B,C,T)).
// chunked = 1: stash is [B][C/4][T][4] (tcgen05 path); 0: stash is (B,T,C) channel-last.
__global__ void feature_convert_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t B, int T, int C,
                                       int chunked, int to_stash) {
  const int64_t n = B * (int64_t)T * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // i enumerates the public layout: (b, c, t)
    const int t = (int)(i % T);
    const int c = (int)((i / T) % C);
    const int64_t b = i / ((int64_t)T * C);
    const int64_t s = chunked ? ((b * (C / 4) + c / 4) * T + t) * 4 + (c & 3) : (b * T + t) * C + c;
    if (to_stash) out[s] = in[i];
    else out[i] = in[s];
  }
}
}  // namespace

extern "C" int wgg_disc_feature_convert(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* in, float* out, int64_t B,
                                        int32_t T, int32_t C, int to_stash, void* stream) {
  DLayout d;
  if (!ctx) return WGG_EINVAL;
  if (disc_layout(cfg, &d) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "disc: bad config%s");
  if (B <= 0) return WGG_OK;
  feature_convert_kernel<<<ew_blocks(B * T * C), 256, 0, (cudaStream_t)stream>>>(in, out, B, T, C,
                                                                                 disc_use_tc(ctx, d) ? 1 : 0, to_stash);
  WGG_CHECK_LAUNCH(ctx, "feature_convert_kernel");
  return WGG_OK;
}

extern "C" int wgg_transpose_tc(wgg_ctx* ctx, const float* in, float* out, int64_t B, int32_t T, int32_t C,
                                void* stream) {
  if (!ctx) return WGG_EINVAL;
  if (B <= 0) return WGG_OK;
  if (B > 65535) return wgg_fail(ctx, WGG_EINVAL, "transpose_tc: batch too large for one launch%s");
  dim3 grid((T + 31) / 32, (C + 31) / 32, (unsigned)B);
  transpose_tc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, B, T, C);
  WGG_CHECK_LAUNCH(ctx, "transpose_tc_kernel");
  return WGG_OK;
}

// Generator = word-prototype-conditioned 4-layer BiLSTM + tanh(Linear) head, forward and backward.
// Replaces Generator.forward (src/gan/models.py:125-165: slice/repeat/cat :147-157, nn.LSTM :160,
// tanh(Linear) :163) and its autograd.  LSTM cell equations: nn.LSTM, gate order i,f,g,o, two biases,
// h0 = c0 = 0; reverse direction scans t = T-1..0; layer l>=1 input = concat(h_fwd, h_bwd).
//
// Data layout in HBM (ours, time-major so that one timestep of a sample tile is one contiguous block):
//   x0    [T][B][I0]        layer-0 input  (prototype[:, :, :pd] ++ z)
//   hseq  [T][B][2H]        per layer output (fwd dir in [0,H), reverse dir in [H,2H))
//   gates [2][T][B][4H]     per layer: input projection + biases, overwritten in place by the
//                           post-activation gates (forward) and then by d(pre-activation) (backward)
//   cseq  [2][T][B][H]      per layer cell state
//
// Structure: the time-parallel part of every layer (x_t W_ih^T for all t, both directions) is one batched
// GEMM; only h_{t-1} W_hh^T runs inside the persistent recurrent kernel, which keeps W_hh resident in
// shared memory for all T steps and the cell state in registers (one lane = one sample, one warp = H/8
// hidden units x 4 gates, so the gate epilogue needs no cross-thread traffic).
#include "common.cuh"

namespace {

struct GenLayout {
  int T, H, L, Z, pd, I0, C;
  int64_t layer_off[WGG_MAX_HIDDEN_LAYERS + 1];  // start of layer l's block in the flat params
  int64_t dir_stride[WGG_MAX_HIDDEN_LAYERS];     // floats between direction 0 and 1 of layer l
  int64_t off_whh[WGG_MAX_HIDDEN_LAYERS], off_bih[WGG_MAX_HIDDEN_LAYERS], off_bhh[WGG_MAX_HIDDEN_LAYERS];
  int64_t off_wo, off_bo, total;
  int in_dim(int l) const { return l == 0 ? I0 : 2 * H; }
};

int gen_layout(const wgg_model_cfg* c, GenLayout* g) {
  if (!c || c->gen_num_layers < 1 || c->gen_num_layers > WGG_MAX_HIDDEN_LAYERS) return WGG_EINVAL;
  g->T = c->seq_length; g->H = c->gen_hidden_dim; g->L = c->gen_num_layers; g->Z = c->latent_dim;
  g->C = c->input_dim;
  g->pd = c->prototype_has_time ? c->input_dim : 2;
  g->I0 = g->pd + g->Z;
  int64_t off = 0;
  for (int l = 0; l < g->L; ++l) {
    const int64_t I = g->in_dim(l), H4 = 4 * g->H;
    g->layer_off[l] = off;
    g->off_whh[l] = H4 * I;
    g->off_bih[l] = g->off_whh[l] + H4 * g->H;
    g->off_bhh[l] = g->off_bih[l] + H4;
    g->dir_stride[l] = g->off_bhh[l] + H4;
    off += 2 * g->dir_stride[l];
  }
  g->layer_off[g->L] = off;
  g->off_wo = off;
  off += (int64_t)g->C * 2 * g->H;
  g->off_bo = off;
  off += g->C;
  g->total = off;
  return WGG_OK;
}

// x0[t][b][j] = j < pd ? proto[b][t][j] : z[b][j-pd]              (models.py:147-157)
__global__ void build_x0_kernel(const float* __restrict__ proto, const float* __restrict__ z, float* __restrict__ x0,
                                int T, int64_t B, int C, int pd, int Z) {
  const int I0 = pd + Z;
  const int64_t n = (int64_t)T * B * I0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % I0);
    const int64_t tb = i / I0;
    const int64_t b = tb % B;
    const int t = (int)(tb / B);
    x0[i] = j < pd ? __ldg(proto + (b * T + t) * C + j) : __ldg(z + b * Z + (j - pd));
  }
}

// Input projection of layer 0 in the scaled regime.  x0[t][b] = [proto[b][t][0..pd) | z[b]] (models.py:147-157): the latent
// part does not depend on t, so  gates[d][t][b][n] = zb[d][b][n] + sum_{c < pd} proto[b][t][c] W_ih[d][n][c]  with
// zb = z W_ih[:, pd:]^T + b_ih + b_hh computed once per gesture (a B x Z x 4H GEMM) - a write-bound fp32 pass over the gate
// buffer instead of a K = pd + Z GEMM whose rows (35 floats) no 16-byte access or TMA box can address.
// block = 128 threads over 4-gate-column groups of one direction; blockIdx.z strides over the (t, b) rows.
__global__ void __launch_bounds__(128) xproj0_kernel(const float* __restrict__ proto, const float* __restrict__ zb,
                                                     const float* __restrict__ w, int64_t dir_stride, float* __restrict__ gates,
                                                     int T, int64_t B, int C, int pd, int I0, int H4) {
  const int n = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (n >= H4) return;
  const int d = blockIdx.y;
  const int64_t TB = (int64_t)T * B;
  float wv[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) wv[j][c] = c < pd ? __ldg(w + d * dir_stride + (int64_t)(n + j) * I0 + c) : 0.f;
  const float* zbd = zb + (int64_t)d * B * H4 + n;
  float* gd = gates + (int64_t)d * TB * H4 + n;
  for (int64_t r = blockIdx.z; r < TB; r += gridDim.z) {
    const int64_t b = r % B;
    const int t = (int)(r / B);
    const float* pp = proto + (b * T + t) * C;
    float x[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = c < pd ? __ldg(pp + c) : 0.f;
    float4 o = *reinterpret_cast<const float4*>(zbd + b * H4);
    o.x += x[0] * wv[0][0] + x[1] * wv[0][1] + x[2] * wv[0][2] + x[3] * wv[0][3];
    o.y += x[0] * wv[1][0] + x[1] * wv[1][1] + x[2] * wv[1][2] + x[3] * wv[1][3];
    o.z += x[0] * wv[2][0] + x[1] * wv[2][1] + x[2] * wv[2][2] + x[3] * wv[2][3];
    o.w += x[0] * wv[3][0] + x[1] * wv[3][1] + x[2] * wv[3][2] + x[3] * wv[3][3];
    *reinterpret_cast<float4*>(gd + r * H4) = o;
  }
}

// The same projection in the chunked gate-buffer order of the persistent H = 128 recurrence (gemm_tc.cu:
// [dir][t][tile of 128 gestures][4H / 4][128 rows][4 floats], thread = gesture, so every access below is a coalesced
// 16-byte access).  zb_chunk_kernel: the per-gesture latent term + biases; xproj0_chunk_kernel: + the prototype term.
__global__ void __launch_bounds__(128) zb_chunk_kernel(const float* __restrict__ z, const float* __restrict__ w, int64_t dir_stride,
                                                       int64_t off_bih, int64_t off_bhh, float* __restrict__ zbc, int64_t B, int Z,
                                                       int pd, int I0, int H4) {
  const int tile = blockIdx.x, d = blockIdx.y, rl = threadIdx.x;
  const int64_t b = (int64_t)tile * 128 + rl;
  float zr[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) zr[k] = (k < Z && b < B) ? __ldg(z + b * Z + k) : 0.f;
  const float* wd = w + d * dir_stride;
  float* out = zbc + (((int64_t)d * gridDim.x + tile) * (H4 / 4) * 128 + rl) * 4;
  for (int n4 = blockIdx.z; n4 < H4 / 4; n4 += gridDim.z) {  // column groups spread over blockIdx.z: enough blocks to fill the SMs
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = 4 * n4 + j;
      const float* wr = wd + (int64_t)n * I0 + pd;
      float acc = __ldg(wd + off_bih + n) + __ldg(wd + off_bhh + n);
#pragma unroll
      for (int k = 0; k < 64; ++k)
        if (k < Z) acc = fmaf(zr[k], __ldg(wr + k), acc);
      o[j] = acc;
    }
    *reinterpret_cast<float4*>(out + (int64_t)n4 * 512) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// block = (tile of 128 gestures, direction, group of 16 gate columns); thread = gesture.  The 16 columns' prototype weights
// and latent terms stay in registers for all T timesteps; the tile's prototype points go through shared memory 32 timesteps
// at a time (coalesced reads of each gesture's contiguous (t, c) run); every store is a coalesced 16-byte store.
constexpr int XP_TS = 32;  // timesteps staged per pass
__global__ void __launch_bounds__(128) xproj0_chunk_kernel(const float* __restrict__ proto, const float* __restrict__ zbc,
                                                           const float* __restrict__ w, int64_t dir_stride,
                                                           float* __restrict__ gates, int T, int64_t B, int C, int pd, int I0,
                                                           int H4) {
  extern __shared__ float s_p[];  // [128 rows][XP_TS * C + 1]
  const int tile = blockIdx.x, d = blockIdx.y, cg = blockIdx.z, rl = threadIdx.x, tiles = gridDim.x;
  const int ld = XP_TS * C + 1;
  const int64_t b = (int64_t)tile * 128 + rl;
  const int64_t b0 = (int64_t)tile * 128;
  const int nrows = B - b0 < 128 ? (int)(B - b0) : 128;
  float wv[16][3], zv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float* wr = w + d * dir_stride + (int64_t)(cg * 16 + j) * I0;
#pragma unroll
    for (int c = 0; c < 3; ++c) wv[j][c] = c < pd ? __ldg(wr + c) : 0.f;
  }
  const float* zb = zbc + ((((int64_t)d * tiles + tile) * (H4 / 4) + cg * 4) * 128 + rl) * 4;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(zb + (int64_t)q * 512);
    zv[4 * q] = v.x; zv[4 * q + 1] = v.y; zv[4 * q + 2] = v.z; zv[4 * q + 3] = v.w;
  }
  for (int t0 = 0; t0 < T; t0 += XP_TS) {
    const int nt = T - t0 < XP_TS ? T - t0 : XP_TS;
    __syncthreads();
    // rows of nt * C contiguous floats: consecutive threads read consecutive floats of one row
    for (int i = threadIdx.x; i < nrows * nt * C; i += 128) {
      const int r = i / (nt * C), k = i % (nt * C);
      s_p[r * ld + k] = __ldg(proto + ((b0 + r) * T + t0) * C + k);
    }
    __syncthreads();
    if (b < B) {
      for (int tt = 0; tt < nt; ++tt) {
        const float* pp = s_p + rl * ld + tt * C;
        const float x0 = pp[0], x1 = pd > 1 ? pp[1] : 0.f, x2 = pd > 2 ? pp[2] : 0.f;
        float* g = gates + (((((int64_t)d * T + t0 + tt) * tiles + tile) * (H4 / 4) + cg * 4) * 128 + rl) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o;
          o.x = zv[4 * q] + x0 * wv[4 * q][0] + x1 * wv[4 * q][1] + x2 * wv[4 * q][2];
          o.y = zv[4 * q + 1] + x0 * wv[4 * q + 1][0] + x1 * wv[4 * q + 1][1] + x2 * wv[4 * q + 1][2];
          o.z = zv[4 * q + 2] + x0 * wv[4 * q + 2][0] + x1 * wv[4 * q + 2][1] + x2 * wv[4 * q + 2][2];
          o.w = zv[4 * q + 3] + x0 * wv[4 * q + 3][0] + x1 * wv[4 * q + 3][1] + x2 * wv[4 * q + 3][2];
          *reinterpret_cast<float4*>(g + (int64_t)q * 512) = o;
        }
      }
    }
  }
}

// Output head of the scaled path: out[b][t][c] = tanh(h[t][b][:] . Wo[c] + bo[c])  (models.py:161-163; fp32).  One warp per
// (t, b) row: a coalesced read of the 2H hidden values, Wo in shared memory, warp-shuffle reduction - a pass over hseq at
// memory speed instead of T batched N = 3 GEMMs.
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ h, const float* __restrict__ wo,
                                                       const float* __restrict__ bo, float* __restrict__ out, int T, int64_t B,
                                                       int K, int C) {
  extern __shared__ float s_wo[];  // [C][K]
  for (int i = threadIdx.x; i < C * K; i += 256) s_wo[i] = __ldg(wo + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t nrows = (int64_t)T * B;
  const int64_t warps = (int64_t)gridDim.x * 8;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < nrows; row += warps) {
    const float* hr = h + row * K;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 4 * lane; k < K; k += 128) {
      const float4 x = *reinterpret_cast<const float4*>(hr + k);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < C) {
          const float4 wv = *reinterpret_cast<const float4*>(s_wo + c * K + k);
          acc[c] += x.x * wv.x + x.y * wv.y + x.z * wv.z + x.w * wv.w;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = warp_sum(acc[c]);
    if (lane < C) {
      const float v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      const int64_t t = row / B, b = row % B;
      out[(b * T + t) * C + lane] = tanhf(v + __ldg(bo + lane));
    }
  }
}

// dpre[t][b][c] = dy[b][t][c] * (1 - y[b][t][c]^2)                (backward of tanh, models.py:163)
__global__ void head_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dpre,
                                int T, int64_t B, int C) {
  const int64_t n = (int64_t)T * B * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t tb = i / C;
    const int64_t b = tb % B;
    const int t = (int)(tb / B);
    const int64_t src = (b * T + t) * C + c;
    const float yy = __ldg(y + src);
    dpre[i] = __ldg(dy + src) * (1.f - yy * yy);
  }
}

// dz[b][j] = sum_t dx0[t][b][pd+j]                                  (backward of repeat+cat)
__global__ void dz_kernel(const float* __restrict__ dx0, float* __restrict__ dz, int T, int64_t B, int I0, int pd,
                          int Z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Z) return;
  const int64_t b = i / Z;
  const int j = (int)(i % Z);
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += __ldg(dx0 + ((int64_t)t * B + b) * I0 + pd + j);
  dz[i] = s;
}

// ---------------------------------------------------------------------------------------------
// persistent recurrent forward: grid (ceil(B/32), 2 directions), 256 threads.
// ---------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(256) lstm_rec_fwd_kernel(float* __restrict__ gates, const float* __restrict__ lp,
                                                           int64_t dir_stride, int64_t off_whh,
                                                           float* __restrict__ hseq, float* __restrict__ cseq,
                                                           int T, int64_t B, int store) {
  constexpr int UPT = H / 8;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;              // [H(k)][H(u)][4(g)]
  float* hs = smem + H * H * 4;  // [2][H][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dir = blockIdx.y;
  const int64_t b = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = b < B;
  const float* __restrict__ whh = lp + dir * dir_stride + off_whh;
  for (int idx = tid; idx < 4 * H * H; idx += 256) {
    const int row = idx / H, k = idx % H;
    const int g = row / H, u = row % H;
    Ws[(k * H + u) * 4 + g] = __ldg(whh + idx);
  }
  for (int idx = tid; idx < 2 * H * 32; idx += 256) hs[idx] = 0.f;
  __syncthreads();

  const int u0 = warp * UPT;
  float c[UPT];
#pragma unroll
  for (int uu = 0; uu < UPT; ++uu) c[uu] = 0.f;
  const int64_t TB = (int64_t)T * B;
  float* gbase = gates + (int64_t)dir * TB * 4 * H;
  float* cbase = cseq ? cseq + (int64_t)dir * TB * H : nullptr;

  float nxt[UPT][4];
  {
    const int t0 = dir ? T - 1 : 0;
    const float* gp = gbase + ((int64_t)t0 * B + b) * 4 * H;
#pragma unroll
    for (int uu = 0; uu < UPT; ++uu)
#pragma unroll
      for (int g = 0; g < 4; ++g) nxt[uu][g] = valid ? gp[g * H + u0 + uu] : 0.f;
  }
  for (int step = 0; step < T; ++step) {
    const int t = dir ? T - 1 - step : step;
    float acc[UPT][4];
#pragma unroll
    for (int uu = 0; uu < UPT; ++uu)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[uu][g] = nxt[uu][g];
    if (step + 1 < T) {  // prefetch the next step's input projection while this step computes
      const int tn = dir ? t - 1 : t + 1;
      const float* gp = gbase + ((int64_t)tn * B + b) * 4 * H;
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu)
#pragma unroll
        for (int g = 0; g < 4; ++g) nxt[uu][g] = valid ? gp[g * H + u0 + uu] : 0.f;
    }
    const float* hcur = hs + (step & 1) * H * 32;
    float* hnext = hs + ((step + 1) & 1) * H * 32;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float hk = hcur[k * 32 + lane];
      const float4* w4 = reinterpret_cast<const float4*>(Ws + (k * H + u0) * 4);
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) {
        const float4 w = w4[uu];
        acc[uu][0] = fmaf(hk, w.x, acc[uu][0]);
        acc[uu][1] = fmaf(hk, w.y, acc[uu][1]);
        acc[uu][2] = fmaf(hk, w.z, acc[uu][2]);
        acc[uu][3] = fmaf(hk, w.w, acc[uu][3]);
      }
    }
    float* gp = gbase + ((int64_t)t * B + b) * 4 * H;
    float* hp = hseq + ((int64_t)t * B + b) * 2 * H + dir * H;
#pragma unroll
    for (int uu = 0; uu < UPT; ++uu) {
      const float ig = sigmoid_f(acc[uu][0]);
      const float fg = sigmoid_f(acc[uu][1]);
      const float gg = tanhf(acc[uu][2]);
      const float og = sigmoid_f(acc[uu][3]);
      c[uu] = fg * c[uu] + ig * gg;
      const float h = og * tanhf(c[uu]);
      hnext[(u0 + uu) * 32 + lane] = h;
      if (valid) {
        hp[u0 + uu] = h;
        if (store) {
          gp[0 * H + u0 + uu] = ig;
          gp[1 * H + u0 + uu] = fg;
          gp[2 * H + u0 + uu] = gg;
          gp[3 * H + u0 + uu] = og;
          cbase[((int64_t)t * B + b) * H + u0 + uu] = c[uu];
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// persistent recurrent backward (BPTT): same decomposition; writes d(pre-activation) over `gates`.
// ---------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(256) lstm_rec_bwd_kernel(float* __restrict__ gates, const float* __restrict__ cseq,
                                                           const float* __restrict__ lp, int64_t dir_stride,
                                                           int64_t off_whh, const float* __restrict__ dh_out, int T,
                                                           int64_t B) {
  constexpr int UPT = H / 8;
  extern __shared__ __align__(16) float smem[];
  float* Wb = smem;                   // [4H][H]
  float* das = Wb + 4 * H * H;        // [4H][32]
  float* dhs = das + 4 * H * 32;      // [H][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dir = blockIdx.y;
  const int64_t b = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = b < B;
  const float* __restrict__ whh = lp + dir * dir_stride + off_whh;
  for (int idx = tid; idx < 4 * H * H; idx += 256) Wb[idx] = __ldg(whh + idx);
  for (int idx = tid; idx < H * 32; idx += 256) dhs[idx] = 0.f;
  __syncthreads();
  const int u0 = warp * UPT;
  const int64_t TB = (int64_t)T * B;
  float* gbase = gates + (int64_t)dir * TB * 4 * H;
  const float* cbase = cseq + (int64_t)dir * TB * H;
  float dc[UPT];
#pragma unroll
  for (int uu = 0; uu < UPT; ++uu) dc[uu] = 0.f;

  for (int step = T - 1; step >= 0; --step) {
    const int t = dir ? T - 1 - step : step;
    const int tp = dir ? t + 1 : t - 1;  // the timestep processed before t in the forward scan
    float* gp = gbase + ((int64_t)t * B + b) * 4 * H;
    const float* cp = cbase + ((int64_t)t * B + b) * H;
    const float* cpp = cbase + ((int64_t)tp * B + b) * H;
    const float* dhp = dh_out + ((int64_t)t * B + b) * 2 * H + dir * H;
#pragma unroll
    for (int uu = 0; uu < UPT; ++uu) {
      const int u = u0 + uu;
      float da_i = 0.f, da_f = 0.f, da_g = 0.f, da_o = 0.f;
      if (valid) {
        const float ig = gp[0 * H + u], fg = gp[1 * H + u], gg = gp[2 * H + u], og = gp[3 * H + u];
        const float cc = cp[u];
        const float cprev = step > 0 ? cpp[u] : 0.f;
        const float tc = tanhf(cc);
        const float dh = dhp[u] + dhs[u * 32 + lane];
        const float d_o = dh * tc;
        const float dct = dc[uu] + dh * og * (1.f - tc * tc);
        da_i = dct * gg * ig * (1.f - ig);
        da_f = dct * cprev * fg * (1.f - fg);
        da_g = dct * ig * (1.f - gg * gg);
        da_o = d_o * og * (1.f - og);
        dc[uu] = dct * fg;
        gp[0 * H + u] = da_i;
        gp[1 * H + u] = da_f;
        gp[2 * H + u] = da_g;
        gp[3 * H + u] = da_o;
      }
      das[(0 * H + u) * 32 + lane] = da_i;
      das[(1 * H + u) * 32 + lane] = da_f;
      das[(2 * H + u) * 32 + lane] = da_g;
      das[(3 * H + u) * 32 + lane] = da_o;
    }
    __syncthreads();
    if (step > 0) {
      float acc[UPT];
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) acc[uu] = 0.f;
#pragma unroll 4
      for (int j = 0; j < 4 * H; ++j) {
        const float d = das[j * 32 + lane];
        const float* w = Wb + j * H + u0;
#pragma unroll
        for (int uu = 0; uu < UPT; ++uu) acc[uu] = fmaf(d, w[uu], acc[uu]);
      }
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) dhs[(u0 + uu) * 32 + lane] = acc[uu];
    }
    __syncthreads();
  }
}

template <int H>
int rec_fwd_launch_t(wgg_ctx* ctx, float* gates, const float* lp, int64_t dir_stride, int64_t off_whh, float* hseq,
                     float* cseq, int T, int64_t B, int store, cudaStream_t st) {
  const size_t smem = (size_t)(H * H * 4 + 2 * H * 32) * sizeof(float);
  if (!wgg_smem_ok(ctx, lstm_rec_fwd_kernel<H>, smem))
    return wgg_fail(ctx, WGG_ECUDA, "lstm_rec_fwd_kernel: cannot reserve shared memory%s");
  dim3 grid((unsigned)cdiv64(B, 32), 2);
  // algorithmic work: 2 dirs x T steps x B x (8 H^2 MAC-flops + gate math); bytes: xproj in, gates/c/h out
  ProfScope prof(ctx, "lstm_rec_fwd_kernel", st, 2.0 * T * (double)B * 8.0 * H * H,
                 2.0 * T * (double)B * 4.0 * (4 * H + (store ? 5 * H : 0) + H));
  lstm_rec_fwd_kernel<H><<<grid, 256, smem, st>>>(gates, lp, dir_stride, off_whh, hseq, cseq, T, B, store);
  WGG_CHECK_LAUNCH(ctx, "lstm_rec_fwd_kernel");
  return WGG_OK;
}

template <int H>
int rec_bwd_launch_t(wgg_ctx* ctx, float* gates, const float* cseq, const float* lp, int64_t dir_stride,
                     int64_t off_whh, const float* dh_out, int T, int64_t B, cudaStream_t st) {
  const size_t smem = (size_t)(4 * H * H + 4 * H * 32 + H * 32) * sizeof(float);
  if (!wgg_smem_ok(ctx, lstm_rec_bwd_kernel<H>, smem))
    return wgg_fail(ctx, WGG_ECUDA, "lstm_rec_bwd_kernel: cannot reserve shared memory%s");
  dim3 grid((unsigned)cdiv64(B, 32), 2);
  ProfScope prof(ctx, "lstm_rec_bwd_kernel", st, 2.0 * T * (double)B * 8.0 * H * H,
                 2.0 * T * (double)B * 4.0 * (4 * H + 4 * H + 2 * H + H));
  lstm_rec_bwd_kernel<H><<<grid, 256, smem, st>>>(gates, cseq, lp, dir_stride, off_whh, dh_out, T, B);
  WGG_CHECK_LAUNCH(ctx, "lstm_rec_bwd_kernel");
  return WGG_OK;
}


// ---------------------------------------------------------------------------------------------
// Any hidden size (the scaled-model regime, BASELINE configs[3]: H = 128 ... 1024, T = 256; config.py:22 is a free
// knob in the reference).  W_hh no longer fits shared memory, so the recurrence runs step by step: per timestep one
// batched (both directions) GEMM  gates[d][t] += h[d][t_prev] W_hh[d]^T  on the contraction engine (TF32 tensor cores in
// the tensor-core math modes, fp32 FMA otherwise) followed by one cell kernel; backward mirrors it with
// dh_rec[d] = da[d][t] W_hh[d].  Same HBM layouts and the same stash conventions as the persistent kernels above
// (gates hold the input projection, then the activated gates, then d(pre-activation)), so every other GEMM of the
// layer (input projection, dW_ih, dW_hh, db, dx) is shared with them.
// ---------------------------------------------------------------------------------------------
__global__ void lstm_cell_fwd_kernel(float* __restrict__ gates, float* __restrict__ cseq, float* __restrict__ cstate,
                                     float* __restrict__ hseq, int T, int64_t B, int H, int step, int store) {
  const int64_t n = 2 * B * H;
  const int64_t TB = (int64_t)T * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % H);
    const int64_t b = (i / H) % B;
    const int dir = (int)(i / ((int64_t)H * B));
    const int t = dir ? T - 1 - step : step;
    const int tp = dir ? t + 1 : t - 1;
    float* gp = gates + (((int64_t)dir * T + t) * B + b) * 4 * H;
    const float ig = sigmoid_f(gp[u]);
    const float fg = sigmoid_f(gp[H + u]);
    const float gg = tanhf(gp[2 * H + u]);
    const float og = sigmoid_f(gp[3 * H + u]);
    float cprev = 0.f;
    if (step > 0) cprev = store ? cseq[(int64_t)dir * TB * H + ((int64_t)tp * B + b) * H + u] : cstate[i];
    const float c = fg * cprev + ig * gg;
    hseq[((int64_t)t * B + b) * 2 * H + dir * H + u] = og * tanhf(c);
    if (store) {
      gp[u] = ig; gp[H + u] = fg; gp[2 * H + u] = gg; gp[3 * H + u] = og;
      cseq[(int64_t)dir * TB * H + ((int64_t)t * B + b) * H + u] = c;
    } else {
      cstate[i] = c;
    }
  }
}

// da over the activated gates; dc carried in `dcs` [2][B][H]; dh_rec [2][B][H] is the previous launch's GEMM result
__global__ void lstm_cell_bwd_kernel(float* __restrict__ gates, const float* __restrict__ cseq, const float* __restrict__ dh_out,
                                     const float* __restrict__ dhrec, float* __restrict__ dcs, int T, int64_t B, int H,
                                     int step) {
  const int64_t n = 2 * B * H;
  const int64_t TB = (int64_t)T * B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % H);
    const int64_t b = (i / H) % B;
    const int dir = (int)(i / ((int64_t)H * B));
    const int t = dir ? T - 1 - step : step;
    const int tp = dir ? t + 1 : t - 1;
    float* gp = gates + (((int64_t)dir * T + t) * B + b) * 4 * H;
    const float* cb = cseq + (int64_t)dir * TB * H;
    const float ig = gp[u], fg = gp[H + u], gg = gp[2 * H + u], og = gp[3 * H + u];
    const float cc = cb[((int64_t)t * B + b) * H + u];
    const float cprev = step > 0 ? cb[((int64_t)tp * B + b) * H + u] : 0.f;
    const float tc = tanhf(cc);
    const float dh = dh_out[((int64_t)t * B + b) * 2 * H + dir * H + u] + (step < T - 1 ? dhrec[i] : 0.f);
    const float d_o = dh * tc;
    const float dct = (step < T - 1 ? dcs[i] : 0.f) + dh * og * (1.f - tc * tc);
    gp[u] = dct * gg * ig * (1.f - ig);
    gp[H + u] = dct * cprev * fg * (1.f - fg);
    gp[2 * H + u] = dct * ig * (1.f - gg * gg);
    gp[3 * H + u] = d_o * og * (1.f - og);
    dcs[i] = dct * fg;
  }
}

bool rec_has_persistent_kernel(int H) { return H == 8 || H == 16 || H == 32 || H == 48 || H == 64; }

// scratch (floats) of the step-by-step path: forward without a stash keeps the cell state [2][B][H];
// backward keeps dh_rec and dc [2][B][H] each
int64_t rec_generic_scratch_floats(int H, int64_t B, int backward) {
  if (rec_has_persistent_kernel(H)) return 0;
  // backward: dh_rec | dc, or (fused tcgen05 step kernels) dc | the W_hh^T images
  // (the carried dc of the chunked BPTT is padded to whole 128-gesture tiles)
  return backward ? 4 * ((B + 127) / 128 * 128) * (int64_t)H + 8 * (int64_t)H * H + 64 : 2 * ((B + 127) / 128 * 128) * (int64_t)H;
}

// K-major operand images of the tcgen05 weight / input-gradient GEMMs of the step-by-step path (backward only):
// da^T [2][4H][T B] | two ping-pong h^T / input^T buffers [2H][T B] | W_ih^T [2][maxI][4H] | row-major copy of da
// [2][T B][4H] (chunked stash only: the input-gradient GEMM and the bias sums read it)  (+ alignment slack)
int64_t wgrad_tc_scratch_floats(int H, int T, int64_t B, int64_t maxI) {
  if (rec_has_persistent_kernel(H)) return 0;
  const int64_t TB = (int64_t)T * B;
  const int64_t nblk = (int64_t)T * ((B + 127) / 128);  // (t, tile) blocks of the chunked layout
  // + first-stage partials [planes][2][T * tiles][4H] written by unchunk_da_kernel (bias sums; layer 0: + 3 prototype planes)
  return TB * 8 * H + 2 * TB * 2 * H + 2 * maxI * 4 * H + TB * 8 * H + 4 * 2 * nblk * 4 * H + 32;  // 4 planes (layer 0)
}

int rec_fwd_generic(wgg_ctx* ctx, int H, float* gates, const float* lp, int64_t dir_stride, int64_t off_whh, float* hseq,
                    float* cseq, float* cstate, int T, int64_t B, int store, cudaStream_t st) {
  const int64_t TB = (int64_t)T * B;
  const int H4 = 4 * H;
  // tensor-core math modes: H = 128 has a persistent kernel (one launch per layer); otherwise one fused launch per
  // timestep (tcgen05 recurrent product + cell update, gemm_tc.cu)
  if (lstm128_persist_usable(ctx, H, gates, hseq, lp, off_whh, dir_stride) && lstm128_persist_rowmajor())
    return lstm128_persist_forward(ctx, gates, lp, dir_stride, off_whh, hseq, cseq, T, B, store, 0, st);
  if (lstm_step_tc_usable(ctx, H, gates, hseq, lp, off_whh, dir_stride))
    return lstm_step_tc_forward(ctx, H, gates, lp, dir_stride, off_whh, hseq, cseq, cstate, T, B, store, 0, st);
  for (int step = 0; step < T; ++step) {
    if (step > 0) {
      // direction 0 consumes h at t - 1 and writes gates at t = step; direction 1 consumes h at t + 1, writes t = T-1-step
      const float* a0 = hseq + (int64_t)(step - 1) * B * 2 * H;
      const float* a1 = hseq + (int64_t)(T - step) * B * 2 * H + H;
      float* c0 = gates + (int64_t)step * B * H4;
      float* c1 = gates + TB * H4 + (int64_t)(T - 1 - step) * B * H4;
      GemmP p;
      p.tag = "gemm_kernel/lstm_rec_step";
      p.A = a0; p.M = B; p.K = H; p.sam = 2 * H; p.sak = 1;
      p.B = lp + off_whh; p.N = H4; p.sbk = 1; p.sbn = H;
      p.C = c0; p.scm = H4; p.scn = 1; p.accumulate = 1;
      p.nbatch = 2; p.bsA = a1 - a0; p.bsB = dir_stride; p.bsC = c1 - c0;
      WGG_TRY(gemm_launch(ctx, p, st));
    }
    lstm_cell_fwd_kernel<<<ew_blocks(2 * B * H), 256, 0, st>>>(gates, cseq, cstate, hseq, T, B, H, step, store);
    WGG_CHECK_LAUNCH(ctx, "lstm_cell_fwd_kernel");
  }
  return WGG_OK;
}

int rec_bwd_generic(wgg_ctx* ctx, int H, float* gates, const float* cseq, const float* lp, int64_t dir_stride,
                    int64_t off_whh, const float* dh_out, float* scratch, int T, int64_t B, cudaStream_t st) {
  const int64_t TB = (int64_t)T * B;
  const int H4 = 4 * H;
  if (lstm_step_tc_usable(ctx, H, gates, dh_out, lp, off_whh, dir_stride))
    return lstm_step_tc_backward(ctx, H, gates, cseq, lp, dir_stride, off_whh, dh_out, scratch, T, B, 0, st);
  float* dhrec = scratch;
  float* dcs = scratch + 2 * B * H;
  for (int step = T - 1; step >= 0; --step) {
    lstm_cell_bwd_kernel<<<ew_blocks(2 * B * H), 256, 0, st>>>(gates, cseq, dh_out, dhrec, dcs, T, B, H, step);
    WGG_CHECK_LAUNCH(ctx, "lstm_cell_bwd_kernel");
    if (step > 0) {
      const float* a0 = gates + (int64_t)step * B * H4;
      const float* a1 = gates + TB * H4 + (int64_t)(T - 1 - step) * B * H4;
      GemmP p;  // dh_rec[d] (B x H) = da[d][t] (B x 4H) * W_hh[d] (4H x H)
      p.tag = "gemm_kernel/lstm_rec_step_bwd";
      p.A = a0; p.M = B; p.K = H4; p.sam = H4; p.sak = 1;
      p.B = lp + off_whh; p.N = H; p.sbk = H; p.sbn = 1;
      p.C = dhrec; p.scm = H; p.scn = 1;
      p.nbatch = 2; p.bsA = a1 - a0; p.bsB = dir_stride; p.bsC = B * H;
      WGG_TRY(gemm_launch(ctx, p, st));
    }
  }
  return WGG_OK;
}

#define WGG_DISPATCH_H(H, CALL)                                                                         \
  switch (H) {                                                                                          \
    case 8: { constexpr int HH = 8; return CALL; }                                                      \
    case 16: { constexpr int HH = 16; return CALL; }                                                    \
    case 32: { constexpr int HH = 32; return CALL; }                                                    \
    case 48: { constexpr int HH = 48; return CALL; }                                                    \
    case 64: { constexpr int HH = 64; return CALL; }                                                    \
    default:                                                                                            \
      return wgg_fail(ctx, WGG_EUNSUPPORTED, "gen_hidden_dim=%s%lld has no compiled recurrent kernel (8,16,32,48,64)", \
                      "", (long long)(H));                                                              \
  }

int rec_fwd_launch(wgg_ctx* ctx, int H, float* gates, const float* lp, int64_t dir_stride, int64_t off_whh,
                   float* hseq, float* cseq, float* scratch, int T, int64_t B, int store, cudaStream_t st) {
  if (!rec_has_persistent_kernel(H))
    return rec_fwd_generic(ctx, H, gates, lp, dir_stride, off_whh, hseq, cseq, scratch, T, B, store, st);
  WGG_DISPATCH_H(H, (rec_fwd_launch_t<HH>(ctx, gates, lp, dir_stride, off_whh, hseq, cseq, T, B, store, st)));
}

int rec_bwd_launch(wgg_ctx* ctx, int H, float* gates, const float* cseq, const float* lp, int64_t dir_stride,
                   int64_t off_whh, const float* dh_out, float* scratch, int T, int64_t B, cudaStream_t st) {
  if (!rec_has_persistent_kernel(H))
    return rec_bwd_generic(ctx, H, gates, cseq, lp, dir_stride, off_whh, dh_out, scratch, T, B, st);
  WGG_DISPATCH_H(H, (rec_bwd_launch_t<HH>(ctx, gates, cseq, lp, dir_stride, off_whh, dh_out, T, B, st)));
}

inline int ew_grid(int64_t n) {
  int64_t g = cdiv64(n, 256);
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

struct StashView {
  float* zb;  // scaled regime: per-gesture latent term of the layer-0 projection
  float* x0;
  float* hseq[WGG_MAX_HIDDEN_LAYERS];
  float* gates[WGG_MAX_HIDDEN_LAYERS];
  float* cseq[WGG_MAX_HIDDEN_LAYERS];
};

inline int64_t pad128(int64_t B) { return (B + 127) / 128 * 128; }
// gate buffer of a no-grad forward: [2][T][B][4H], in the scaled regime with B padded to whole tiles
int64_t fwd_gate_floats(const GenLayout& g, int64_t B) {
  return (int64_t)g.T * (rec_has_persistent_kernel(g.H) ? B : pad128(B)) * 8 * g.H;
}

// second stage of the bias gradients: out[d][col] (+)= sum over the S blocks of colpart[d][s][col], the same into out2.
// block = 32 columns x 8 interleaved segments of the S partials (8 independent loads in flight per thread), fixed order.
__global__ void __launch_bounds__(256) colpart_reduce_kernel(const float* __restrict__ colpart, int S, int C4, float* __restrict__ out,
                                                             float* __restrict__ out2, int64_t bsOut, int64_t ostride) {
  __shared__ float part[8][32];
  const int c = threadIdx.x & 31, seg = threadIdx.x >> 5, d = blockIdx.y;
  const int col = blockIdx.x * 32 + c;
  const float* p = colpart + (int64_t)d * S * C4 + col;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int s = seg;
  for (; s + 56 < S; s += 64) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += p[(int64_t)(s + 8 * j) * C4];
  }
  for (; s < S; s += 8) acc[0] += p[(int64_t)s * C4];
  part[seg][c] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (seg == 0) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += part[k][c];
    out[(int64_t)d * bsOut + col * ostride] += tot;
    if (out2) out2[(int64_t)d * bsOut + col * ostride] += tot;
  }
}

// Output head backward of the scaled path in one pass over hseq (models.py:161-163):
//   dpre[r][k] = dy (1 - y^2);  dh[r][:] = dpre[r] . Wo;  dWo[k][:] += sum_r dpre[r][k] h[r][:];  dbo[k] += sum_r dpre[r][k]
// block: slabs of 64 (t, b) rows; thread = (4-column group of the 2H hidden columns, row phase); per-block partial sums of
// dWo / dbo go to `part` ([blocks][C * K2 + 4]) and are reduced in fixed order by head_bwd_reduce_kernel.
constexpr int HB_ROWS = 64;
__global__ void __launch_bounds__(256) head_bwd_fused_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                             const float* __restrict__ hL, const float* __restrict__ wo,
                                                             float* __restrict__ dh, float* __restrict__ part, int T, int64_t B,
                                                             int K2, int C) {
  extern __shared__ float s_hb[];           // Wo [C][K2] | dpre [HB_ROWS][4] | partial combine [phases][C * K2]
  float* s_wo = s_hb;
  float* s_dp = s_wo + C * K2;
  float* s_pc = s_dp + HB_ROWS * 4;
  const int ncg = K2 / 4;                   // column groups
  const int phases = 256 / ncg > 0 ? 256 / ncg : 1;  // row phases (threads beyond phases * ncg idle in the column loop)
  const int cg = threadIdx.x % ncg, ph = threadIdx.x / ncg;
  const bool colthread = ph < phases;
  for (int i = threadIdx.x; i < C * K2; i += 256) s_wo[i] = __ldg(wo + i);
  float4 accw[4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) accw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t nrows = (int64_t)T * B;
  for (int64_t r0 = (int64_t)blockIdx.x * HB_ROWS; r0 < nrows; r0 += (int64_t)gridDim.x * HB_ROWS) {
    __syncthreads();
    if (threadIdx.x < HB_ROWS * 4) {
      const int rr = threadIdx.x >> 2, k = threadIdx.x & 3;
      const int64_t r = r0 + rr;
      float v = 0.f;
      if (r < nrows && k < C) {
        const int64_t t = r / B, b = r % B;
        const int64_t src = (b * T + t) * C + k;
        const float yy = __ldg(y + src);
        v = __ldg(dy + src) * (1.f - yy * yy);
      }
      s_dp[threadIdx.x] = v;
    }
    __syncthreads();
    if (colthread) {
      for (int rr = ph; rr < HB_ROWS; rr += phases) {
        const int64_t r = r0 + rr;
        if (r >= nrows) break;
        const float4 h4 = *reinterpret_cast<const float4*>(hL + r * K2 + 4 * cg);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < C) {
            const float d = s_dp[rr * 4 + k];
            const float4 w = *reinterpret_cast<const float4*>(s_wo + k * K2 + 4 * cg);
            o.x += d * w.x; o.y += d * w.y; o.z += d * w.z; o.w += d * w.w;
            accw[k].x += d * h4.x; accw[k].y += d * h4.y; accw[k].z += d * h4.z; accw[k].w += d * h4.w;
            if (cg == 0) accb[k] += d;
          }
        }
        *reinterpret_cast<float4*>(dh + r * K2 + 4 * cg) = o;
      }
    }
  }
  // combine the row phases in fixed order, one partial row per block
  __syncthreads();
  if (colthread)
    for (int k = 0; k < C; ++k) *reinterpret_cast<float4*>(s_pc + ((int64_t)ph * C + k) * K2 + 4 * cg) = accw[k];
  __shared__ float s_b[64][4];
  if (colthread && cg == 0)
    for (int k = 0; k < 4; ++k) s_b[ph][k] = accb[k];
  __syncthreads();
  float* prow = part + (int64_t)blockIdx.x * (C * K2 + 4);
  for (int i = threadIdx.x; i < C * K2; i += 256) {
    float tot = 0.f;
    for (int q = 0; q < phases; ++q) tot += s_pc[(int64_t)q * C * K2 + i];
    prow[i] = tot;
  }
  if (threadIdx.x < 4) {
    float tot = 0.f;
    for (int q = 0; q < phases; ++q) tot += s_b[q][threadIdx.x];
    prow[C * K2 + threadIdx.x] = tot;
  }
}

// dWo (C x K2, contiguous) += sum over the blocks' partials; dbo (C) likewise
__global__ void __launch_bounds__(256) head_bwd_reduce_kernel(const float* __restrict__ part, int nblocks, int K2, int C,
                                                              float* __restrict__ dwo, float* __restrict__ dbo) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int n = C * K2 + 4, nw = C * K2;
  if (i >= n || (i >= nw && i - nw >= C)) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int s = 0;
  for (; s + 3 < nblocks; s += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += part[(int64_t)(s + j) * n + i];
  }
  for (; s < nblocks; ++s) acc[0] += part[(int64_t)s * n + i];
  const float tot = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  if (i < nw) dwo[i] += tot;
  else dbo[i - nw] += tot;
}

// Stash of a grad-carrying forward: x0 | hseq[l] | gates[l] | cseq[l] (| zb).  In the scaled regime the gate and c buffers
// are padded to whole 128-gesture tiles (the chunked order of the persistent H = 128 kernels needs whole tiles; the row-major
// order simply leaves the tail unused) and the per-gesture latent term of the layer-0 projection (zb) lives here too.
int64_t stash_floats(const GenLayout& g, int64_t B) {
  const int64_t TB = (int64_t)g.T * B;
  if (rec_has_persistent_kernel(g.H)) return TB * (g.I0 + (int64_t)g.L * 12 * g.H);
  const int64_t TBp = (int64_t)g.T * pad128(B);
  return TB * (g.I0 + (int64_t)g.L * 2 * g.H) + TBp * (int64_t)g.L * 10 * g.H + 8 * pad128(B) * (int64_t)g.H + 16;
}

void stash_view(const GenLayout& g, int64_t B, float* s, StashView* v) {
  const int64_t TB = (int64_t)g.T * B;
  const int64_t TBg = rec_has_persistent_kernel(g.H) ? TB : (int64_t)g.T * pad128(B);
  v->x0 = s;
  s += TB * g.I0;
  for (int l = 0; l < g.L; ++l) { v->hseq[l] = s; s += TB * 2 * g.H; }
  for (int l = 0; l < g.L; ++l) { v->gates[l] = s; s += TBg * 8 * g.H; }
  for (int l = 0; l < g.L; ++l) { v->cseq[l] = s; s += TBg * 2 * g.H; }
  if (reinterpret_cast<uintptr_t>(s) & 15) s += 4 - ((reinterpret_cast<uintptr_t>(s) & 15) >> 2);
  v->zb = s;
}

// [dir][t][tile][columns / 4][128][4] (chunked da, TF32-rounded by BPTT) -> daT [dir][column][T B] (K-major image for the weight
// gradients) and da_rm [dir][T B][columns] (row-major copy for the input-gradient GEMM and the bias sums).
// block = (t * tiles + tile, group of 32 columns, dir); 256 threads; both outputs written in 128-byte runs.
// colpart [dir][t * tiles + tile][columns]: the block's column sums over its (valid) rows - first stage of the bias
// gradients db_ih = db_hh = sum over (t, b) of da, reduced in fixed order by reduce_partials afterwards.
// Layer 0 (x0 != nullptr: the stashed [T][B][I0] input, whose first pd columns are the prototype point): three more planes
// colpart[1 + k] = sum over the rows of da * prototype column k - first stage of dW_ih[:, k], k < pd <= 3 (the latent columns of
// dW_ih come from the time-summed da, sum_t_chunk_kernel: x0's latent part does not depend on t).  Plane stride 2 * blocks * C4.
__global__ void __launch_bounds__(256) unchunk_da_kernel(const float* __restrict__ dac, float* __restrict__ daT,
                                                         float* __restrict__ da_rm, float* __restrict__ colpart, int T, int64_t B,
                                                         int C4, const float* __restrict__ x0, int I0, int pd) {
  __shared__ float tile[128][33];                   // [row][column]: source of the row-major copy
  __shared__ __align__(16) float tileT[32][132];    // [column][row]: source of the K-major image
  __shared__ float ps[128][3];                      // layer 0: the rows' prototype point at this timestep
  const int tiles = (int)((B + 127) / 128);
  const int t = blockIdx.x / tiles, tl = blockIdx.x % tiles, cg = blockIdx.y, d = blockIdx.z;
  const int64_t TB = (int64_t)T * B;
  const float* src = dac + ((((int64_t)d * T + t) * tiles + tl) * (C4 / 4) + cg * 8) * 512;
  for (int i = threadIdx.x; i < 8 * 128; i += 256) {  // 8 chunks x 128 rows of 16 bytes, contiguous
    const float4 v = *reinterpret_cast<const float4*>(src + (int64_t)i * 4);
    const int ch = i >> 7, r = i & 127;
    tile[r][4 * ch] = v.x; tile[r][4 * ch + 1] = v.y; tile[r][4 * ch + 2] = v.z; tile[r][4 * ch + 3] = v.w;
    tileT[4 * ch][r] = v.x; tileT[4 * ch + 1][r] = v.y; tileT[4 * ch + 2][r] = v.z; tileT[4 * ch + 3][r] = v.w;
  }
  const int64_t b0 = (int64_t)tl * 128;
  const int nrows = B - b0 < 128 ? (int)(B - b0) : 128;
  const int64_t m0 = (int64_t)t * B + b0;
  if (x0)
    for (int i = threadIdx.x; i < 128 * 3; i += 256) {
      const int r = i / 3, k = i % 3;
      ps[r][k] = (r < nrows && k < pd) ? __ldg(x0 + (m0 + r) * I0 + k) : 0.f;
    }
  __syncthreads();
  if ((B & 3) == 0) {  // 16-byte stores (T B and the tile offsets are multiples of 4)
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {  // daT: column-major runs of 128 rows, 4 rows per store
      const int c = i >> 5, r = (i & 31) * 4;
      if (r < nrows)
        *reinterpret_cast<float4*>(daT + ((int64_t)d * C4 + cg * 32 + c) * TB + m0 + r) =
            *reinterpret_cast<const float4*>(&tileT[c][r]);
    }
  } else {
    for (int i = threadIdx.x; i < 32 * 128; i += 256) {
      const int c = i >> 7, r = i & 127;
      if (r < nrows) daT[((int64_t)d * C4 + cg * 32 + c) * TB + m0 + r] = tileT[c][r];
    }
  }
  for (int i = threadIdx.x; da_rm && i < 8 * 128; i += 256) {  // da_rm: rows of 32 columns, 4 columns per store
    const int r = i >> 3, c = (i & 7) * 4;
    if (r < nrows)
      *reinterpret_cast<float4*>(da_rm + ((int64_t)d * TB + m0 + r) * C4 + cg * 32 + c) =
          make_float4(tile[r][c], tile[r][c + 1], tile[r][c + 2], tile[r][c + 3]);
  }
  {  // column sums: thread = (column, segment of 16 rows); the eight segment sums are combined in fixed order
    const int c = threadIdx.x & 31, seg = threadIdx.x >> 5;
    float sum = 0.f, sp[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const int r = seg * 16 + rr;
      if (r < nrows) {
        const float v = tile[r][c];
        sum += v;
        if (x0) { sp[0] += v * ps[r][0]; sp[1] += v * ps[r][1]; sp[2] += v * ps[r][2]; }
      }
    }
    __syncthreads();           // every read of tileT above is done: reuse its first rows as scratch
    tileT[seg][c] = sum;
    if (x0) { tileT[8 + seg][c] = sp[0]; tileT[16 + seg][c] = sp[1]; tileT[24 + seg][c] = sp[2]; }
    __syncthreads();
    const int planes = x0 ? 4 : 1;
    if (seg < planes) {        // warp `seg` combines plane `seg` in fixed order
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += tileT[8 * seg + k][c];
      colpart[(((int64_t)seg * 2 + d) * gridDim.x + blockIdx.x) * C4 + cg * 32 + c] = tot;
    }
  }
}

// S[d][b][col] = sum over t of da[d][t][b][col] (chunked da -> row-major S): the latent columns of layer 0's input and the
// latent-code gradient only see the time-summed da (x0's latent part is the same at every timestep).
// block = 128 gestures of one tile x one 4-column chunk; coalesced 16-byte loads.
__global__ void __launch_bounds__(128) sum_t_chunk_kernel(const float* __restrict__ dac, float* __restrict__ S, int T, int64_t B,
                                                          int C4) {
  const int tl = blockIdx.x, q = blockIdx.y, d = blockIdx.z, rl = threadIdx.x, tiles = gridDim.x;
  const int64_t tstride = (int64_t)tiles * (C4 / 4) * 512;
  const float* p = dac + (((int64_t)d * T * tiles + tl) * (C4 / 4) + q) * 512 + rl * 4;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  int t = 0;
  for (; t + 1 < T; t += 2) {
    const float4 v0 = *reinterpret_cast<const float4*>(p + (int64_t)t * tstride);
    const float4 v1 = *reinterpret_cast<const float4*>(p + (int64_t)(t + 1) * tstride);
    a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
  }
  if (t < T) {
    const float4 v0 = *reinterpret_cast<const float4*>(p + (int64_t)t * tstride);
    a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
  }
  const int64_t b = (int64_t)tl * 128 + rl;
  if (b < B) *reinterpret_cast<float4*>(S + ((int64_t)d * B + b) * C4 + 4 * q) = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
}

}  // namespace

extern "C" int64_t wgg_generator_param_floats(const wgg_model_cfg* cfg) {
  GenLayout g;
  if (gen_layout(cfg, &g) != WGG_OK) return -1;
  return g.total;
}

extern "C" int64_t wgg_generator_stash_floats(const wgg_model_cfg* cfg, int64_t B) {
  GenLayout g;
  if (gen_layout(cfg, &g) != WGG_OK) return -1;
  const int64_t a = stash_floats(g, B);
  const int64_t b = generator_tc_supported(cfg) ? generator_tc_stash_floats(cfg, B) : 0;
  return a > b ? a : b;  // either math mode fits
}

extern "C" int64_t wgg_generator_workspace_floats(const wgg_model_cfg* cfg, int64_t B, int backward) {
  GenLayout g;
  if (gen_layout(cfg, &g) != WGG_OK) return -1;
  const int64_t TB = (int64_t)g.T * B;
  if (!backward) {  // no-grad forward: x0 + two hseq + gates (FMA path) or the tcgen05 path's buffers
    // scaled regime: the gate buffer is padded to whole 128-gesture tiles (chunked order of the persistent H = 128
    // recurrence) + zb [2][B padded][4H] of the layer-0 input projection (xproj0 kernels)
    const int64_t simt = TB * (g.I0 + 4 * g.H) + fwd_gate_floats(g, B) + rec_generic_scratch_floats(g.H, B, 0) +
                         (rec_has_persistent_kernel(g.H) ? 0 : 8 * pad128(B) * (int64_t)g.H + 8);
    const int64_t tcw = generator_tc_workspace_floats(cfg, B);
    return simt > tcw ? simt : tcw;
  }
  const int64_t maxI = g.I0 > 2 * g.H ? g.I0 : 2 * g.H;
  // dpre | dh | dx | split-K partials | column-sum scratch | (tcgen05 path: its own backward workspace)
  return TB * (g.C + 2 * maxI) + gemm_splitk_ws_floats(4 * g.H, maxI, 2) + colsum_ws_floats(4 * g.H, 2) +
         generator_tc_bwd_workspace_floats(cfg, B) + rec_generic_scratch_floats(g.H, B, 1) +
         wgrad_tc_scratch_floats(g.H, g.T, B, maxI) + 32;  // + alignment slack of the sub-buffers
}

namespace {
// input projection of layer l: gates[d] = in * W_ih[d]^T + b_ih[d] + b_hh[d] (both directions batched); chunked = the gate
// buffer in the order of the persistent H = 128 kernels
GemmP xproj_gemm(const GenLayout& g, int64_t B, int l, const float* in, const float* lp, float* gates, bool chunked) {
  const int64_t TB = (int64_t)g.T * B;
  const int I = g.in_dim(l);
  GemmP p;
  p.tag = "gemm_kernel/lstm_xproj";
  p.A = in; p.M = TB; p.K = I; p.sam = I; p.sak = 1;
  p.B = lp; p.N = 4 * g.H; p.sbk = 1; p.sbn = I;
  p.C = gates; p.scm = 4 * g.H; p.scn = 1;
  p.nbatch = 2; p.bsA = 0; p.bsB = g.dir_stride[l]; p.bsC = TB * 4 * g.H; p.bsBias = g.dir_stride[l];
  p.bias = lp + g.off_bih[l]; p.bias2 = lp + g.off_bhh[l];
  if (chunked) { p.out_chunk = 1; p.chunk_B = B; p.bsC = (int64_t)g.T * pad128(B) * 4 * g.H; }
  return p;
}

// Does layer l run on the chunked gate buffer (scaled regime, tensor-core modes: the persistent H = 128 kernel or the chunked
// variants of the per-timestep kernels)?  The answer depends only on the context's math
// mode, the configuration, the batch size and the alignment of the buffers - the backward pass re-derives the forward's
// answer for a stash from the same call.
bool layer_chunked(wgg_ctx* ctx, const GenLayout& g, int64_t B, int l, const float* in, const float* lp, float* gates,
                   const float* hout, const float* zb) {
  static const bool on = [] { const char* e = getenv("WGG_SCALED_CHUNK"); return !(e && e[0] == '0'); }();
  if (!on || !zb || (reinterpret_cast<uintptr_t>(zb) & 15) || g.Z > 64 || g.pd > 3 || (g.H & 7)) return false;
  if (!lstm_step_tc_usable(ctx, g.H, gates, hout, lp, g.off_whh[l], g.dir_stride[l])) return false;
  if (l == 0) return true;
  return gemm_tc_usable(ctx, xproj_gemm(g, B, l, in, lp, gates, true));
}
// WGG_LSTM128_STASH_CHUNK=0: grad-carrying passes keep the row-major stash and the per-timestep kernels (A/B measurements)
bool stash_chunk_enabled() {
  static const bool on = [] { const char* e = getenv("WGG_LSTM128_STASH_CHUNK"); return !(e && e[0] == '0'); }();
  return on;
}
}  // namespace

extern "C" int wgg_generator_forward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* proto,
                                     const float* z, int64_t B, float* out, float* stash, float* ws,
                                     int64_t ws_floats, void* stream) {
  GenLayout g;
  if (!ctx) return WGG_EINVAL;
  if (gen_layout(cfg, &g) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "generator: bad config%s");
  if (B <= 0) return WGG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t TB = (int64_t)g.T * B;
  if (ctx->math_mode >= 1 && generator_tc_supported(cfg)) {
    // tcgen05 path (lstm_tc.cu): no-grad calls (sampling, the critic loop's generations) and, with a stash, the
    // grad-carrying forward
    if (!ws || ws_floats < generator_tc_workspace_floats(cfg, B))
      return wgg_fail(ctx, WGG_EWORKSPACE, "generator_forward: workspace too small for the tcgen05 path%s");
    return generator_forward_tc(ctx, cfg, params, g.layer_off, g.dir_stride, g.off_whh, g.off_bih, g.off_bhh, g.off_wo,
                                g.off_bo, proto, z, B, out, ws, ws_floats, stash, st);
  }
  StashView sv;
  float* hbuf[2] = {nullptr, nullptr};
  float* gates_ws = nullptr;
  float* rec_scratch = nullptr;  // cell state of the step-by-step recurrence when nothing is stashed
  if (stash) {
    stash_view(g, B, stash, &sv);
  } else {
    if (!ws || ws_floats < wgg_generator_workspace_floats(cfg, B, 0))
      return wgg_fail(ctx, WGG_EWORKSPACE, "generator_forward: workspace too small%s");
    sv.x0 = ws;
    hbuf[0] = ws + TB * g.I0;
    hbuf[1] = hbuf[0] + TB * 2 * g.H;
    gates_ws = hbuf[1] + TB * 2 * g.H;
    rec_scratch = gates_ws + fwd_gate_floats(g, B);
  }
  // scaled regime: layer 0's input projection without materialising x0 (the stash still gets x0: the backward reads it)
  float* zb = nullptr;
  if (!rec_has_persistent_kernel(g.H) && g.pd <= 4 && stash) {
    zb = sv.zb;  // part of the stash: the backward pass must be able to re-derive the layout decision without the workspace
    if ((reinterpret_cast<uintptr_t>(zb) | reinterpret_cast<uintptr_t>(sv.gates[0])) & 15) zb = nullptr;
  } else if (!rec_has_persistent_kernel(g.H) && g.pd <= 4 && ws) {
    const int64_t base = TB * (g.I0 + 4 * g.H) + fwd_gate_floats(g, B) + rec_generic_scratch_floats(g.H, B, 0);
    float* q = ws + base;
    if (reinterpret_cast<uintptr_t>(q) & 15) q += 4 - ((reinterpret_cast<uintptr_t>(q) & 15) >> 2);
    const float* g0 = stash ? sv.gates[0] : gates_ws;
    if (q + 8 * pad128(B) * (int64_t)g.H <= ws + ws_floats && (reinterpret_cast<uintptr_t>(g0) & 15) == 0) zb = q;
  }
  const int64_t tiles = pad128(B) / 128;
  if (stash || !zb) {
    build_x0_kernel<<<ew_grid(TB * g.I0), 256, 0, st>>>(proto, z, sv.x0, g.T, B, g.C, g.pd, g.Z);
    WGG_CHECK_LAUNCH(ctx, "build_x0_kernel");
  }
  const float* in = sv.x0;
  float* hout = nullptr;
  for (int l = 0; l < g.L; ++l) {
    const int I = g.in_dim(l);
    const float* lp = params + g.layer_off[l];
    float* gates = stash ? sv.gates[l] : gates_ws;
    float* cseq = stash ? sv.cseq[l] : nullptr;
    hout = stash ? sv.hseq[l] : hbuf[l & 1];
    // H = 128 in the tensor-core modes: the persistent recurrence on the chunked gate buffer (no-grad passes, and - with c
    // and the activated gates stored in the same order - the grad-carrying pass)
    const bool chunked = (!stash || stash_chunk_enabled()) && layer_chunked(ctx, g, B, l, in, lp, gates, hout, zb);
    const GemmP p = xproj_gemm(g, B, l, in, lp, gates, chunked && l > 0);
    if (l == 0 && zb && chunked) {
      dim3 gz((unsigned)tiles, 2, 16);
      zb_chunk_kernel<<<gz, 128, 0, st>>>(z, lp, g.dir_stride[l], g.off_bih[l], g.off_bhh[l], zb, B, g.Z, g.pd, I, 4 * g.H);
      WGG_CHECK_LAUNCH(ctx, "zb_chunk_kernel");
      dim3 grid((unsigned)tiles, 2, (unsigned)(4 * g.H / 16));
      ProfScope prof(ctx, "xproj0_kernel", st, 2.0 * TB * 8.0 * g.H * g.pd, 4.0 * TB * 8.0 * g.H, "xproj0_kernel");
      const size_t xsm = (size_t)128 * (XP_TS * g.C + 1) * sizeof(float);
      if (xsm > 200 * 1024 || !wgg_smem_ok(ctx, xproj0_chunk_kernel, 200 * 1024))
        return wgg_fail(ctx, WGG_ECUDA, "xproj0_chunk_kernel: cannot reserve shared memory%s");
      xproj0_chunk_kernel<<<grid, 128, xsm, st>>>(proto, zb, lp, g.dir_stride[l], gates, g.T, B, g.C, g.pd, I, 4 * g.H);
      WGG_CHECK_LAUNCH(ctx, "xproj0_chunk_kernel");
    } else if (l == 0 && zb) {
      GemmP q;  // zb[d] (B x 4H) = z * W_ih[d][:, pd:]^T + b_ih[d] + b_hh[d]     (fp32)
      q.tag = "gemm_kernel/lstm_xproj0_z";
      q.A = z; q.M = B; q.K = g.Z; q.sam = g.Z; q.sak = 1;
      q.B = lp + g.pd; q.N = 4 * g.H; q.sbk = 1; q.sbn = I;
      q.C = zb; q.scm = 4 * g.H; q.scn = 1;
      q.nbatch = 2; q.bsA = 0; q.bsB = g.dir_stride[l]; q.bsC = B * 4 * g.H; q.bsBias = g.dir_stride[l];
      q.bias = lp + g.off_bih[l]; q.bias2 = lp + g.off_bhh[l]; q.force_fp32 = 1;
      WGG_TRY(gemm_launch(ctx, q, st));
      int64_t gz = TB < 4096 ? TB : 4096;
      dim3 grid((unsigned)cdiv64(g.H, 128), 2, (unsigned)gz);
      ProfScope prof(ctx, "xproj0_kernel", st, 2.0 * TB * 8.0 * g.H * g.pd, 4.0 * TB * 8.0 * g.H, "xproj0_kernel");
      xproj0_kernel<<<grid, 128, 0, st>>>(proto, zb, lp, g.dir_stride[l], gates, g.T, B, g.C, g.pd, I, 4 * g.H);
      WGG_CHECK_LAUNCH(ctx, "xproj0_kernel");
    } else {
      WGG_TRY(gemm_launch(ctx, p, st));
    }
    if (chunked) {
      if (lstm128_persist_usable(ctx, g.H, gates, hout, lp, g.off_whh[l], g.dir_stride[l]))
        WGG_TRY(lstm128_persist_forward(ctx, gates, lp, g.dir_stride[l], g.off_whh[l], hout, cseq, g.T, B, stash ? 1 : 0, 1, st));
      else
        WGG_TRY(lstm_step_tc_forward(ctx, g.H, gates, lp, g.dir_stride[l], g.off_whh[l], hout, cseq, rec_scratch, g.T, B,
                                     stash ? 1 : 0, 1, st));
      in = hout;
      continue;
    }
    WGG_TRY(rec_fwd_launch(ctx, g.H, gates, lp, g.dir_stride[l], g.off_whh[l], hout, cseq, rec_scratch, g.T, B, stash ? 1 : 0, st));
    in = hout;
  }
  if (!rec_has_persistent_kernel(g.H) && g.C <= 4 && (2 * g.H) % 4 == 0 && (reinterpret_cast<uintptr_t>(hout) & 15) == 0 &&
      (size_t)g.C * 2 * g.H * sizeof(float) <= 48 * 1024) {
    int64_t blocks = cdiv64(TB, 8);
    if (blocks > 16 * (int64_t)ctx->sm_count) blocks = 16 * (int64_t)ctx->sm_count;
    ProfScope prof(ctx, "head_fwd_kernel", st, 2.0 * TB * 2.0 * g.H * g.C, 4.0 * TB * (2.0 * g.H + g.C), "head_fwd_kernel");
    head_fwd_kernel<<<(unsigned)blocks, 256, (size_t)g.C * 2 * g.H * sizeof(float), st>>>(hout, params + g.off_wo, params + g.off_bo,
                                                                                         out, g.T, B, 2 * g.H, g.C);
    WGG_CHECK_LAUNCH(ctx, "head_fwd_kernel");
    return WGG_OK;
  }
  GemmP p;  // out[b][t][:] = tanh(h[t][b][:] * Wo^T + bo), batched over t to transpose (t,b)->(b,t)
  p.tag = "gemm_kernel/head_fwd";
  p.A = hout; p.M = B; p.K = 2 * g.H; p.sam = 2 * g.H; p.sak = 1;
  p.B = params + g.off_wo; p.N = g.C; p.sbk = 1; p.sbn = 2 * g.H;
  p.C = out; p.scm = (int64_t)g.T * g.C; p.scn = 1;
  p.nbatch = g.T; p.bsA = B * 2 * g.H; p.bsB = 0; p.bsC = g.C; p.bsBias = 0;
  p.bias = params + g.off_bo; p.act = ACT_TANH; p.force_fp32 = 1;  // nn.Linear head stays fp32
  return gemm_launch(ctx, p, st);
}

extern "C" int wgg_generator_backward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, int64_t B,
                                      float* stash, const float* out, const float* dout, float* dparams, float* dz,
                                      float* ws, int64_t ws_floats, void* stream) {
  GenLayout g;
  if (!ctx) return WGG_EINVAL;
  if (gen_layout(cfg, &g) != WGG_OK) return wgg_fail(ctx, WGG_EINVAL, "generator: bad config%s");
  if (B <= 0) return WGG_OK;
  if (!stash || !ws || ws_floats < wgg_generator_workspace_floats(cfg, B, 1))
    return wgg_fail(ctx, WGG_EWORKSPACE, "generator_backward: stash/workspace missing or too small%s");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t TB = (int64_t)g.T * B;
  const int H = g.H, H4 = 4 * g.H;
  const int64_t maxI = g.I0 > 2 * H ? g.I0 : 2 * H;
  const bool tcp = ctx->math_mode >= 1 && generator_tc_supported(cfg);
  StashView sv;
  if (!tcp) stash_view(g, B, stash, &sv);
  // every sub-buffer starts on a 16-byte boundary (the kernels of the chunked path make 16-byte accesses to dh and the scratch)
  auto a4 = [](int64_t n) { return (n + 3) & ~(int64_t)3; };
  float* dpre = ws;
  float* dh = dpre + a4(TB * g.C);
  float* dx = dh + a4(TB * maxI);
  float* part = dx + a4(TB * maxI);
  float* csws = part + a4(gemm_splitk_ws_floats(H4, maxI, 2));
  float* tcws = csws + a4(colsum_ws_floats(H4, 2));
  float* rec_scratch = tcws + a4(generator_tc_bwd_workspace_floats(cfg, B));  // dh_rec | dc of the step-by-step recurrence
  // scaled regime, tensor-core modes: weight / input gradients as K-major tcgen05 GEMMs over transposed operand images
  const bool wtc = !tcp && !rec_has_persistent_kernel(H) && lstm_wgrad_tc_usable(ctx, H, B, g.T);
  float* daT = rec_scratch + rec_generic_scratch_floats(H, B, 1);
  if (reinterpret_cast<uintptr_t>(daT) & 15) daT += 4 - ((reinterpret_cast<uintptr_t>(daT) & 15) >> 2);
  float* hT[2] = {daT + TB * 2 * H4, daT + TB * 2 * H4 + TB * 2 * H};  // [2H][T B] each
  float* wihT = hT[1] + TB * 2 * H;                                    // [2][I][4H]
  float* da_rm = wihT + 2 * maxI * H4;                                 // [2][T B][4H] (chunked stash only)
  if (reinterpret_cast<uintptr_t>(da_rm) & 15) da_rm += 4 - ((reinterpret_cast<uintptr_t>(da_rm) & 15) >> 2);
  float* colpart = da_rm + TB * 2 * H4;                                // [2][T * tiles][4H] bias-gradient partials
  int hcur = 0;                                                        // hT[hcur] = transposed output of the current layer
  float* zb_stash = nullptr;
  if (!tcp && !rec_has_persistent_kernel(H) && g.pd <= 4) {
    zb_stash = sv.zb;
    if ((reinterpret_cast<uintptr_t>(zb_stash) | reinterpret_cast<uintptr_t>(sv.gates[0])) & 15) zb_stash = nullptr;
  }

  if (tcp) {
    // head and LSTM stack backward run entirely on the tcgen05 path (fused head, BPTT, dx and dW/db kernels)
    return generator_backward_tc_layers(ctx, cfg, params, dparams, g.layer_off, g.dir_stride, g.off_whh, g.off_bih,
                                        g.off_bhh, B, stash, out, dout, g.off_wo, g.off_bo, dz, tcws,
                                        generator_tc_bwd_workspace_floats(cfg, B), st);
  }
  const float* hL = sv.hseq[g.L - 1];
  const int K2 = 2 * H;
  const int hb_ncg = K2 / 4, hb_ph = 256 / hb_ncg > 0 ? 256 / hb_ncg : 1;
  const size_t hb_smem = ((size_t)g.C * K2 + HB_ROWS * 4 + (size_t)hb_ph * g.C * K2) * sizeof(float);
  // scaled path: the whole head backward in one pass over hseq (the partial rows fit the split-K scratch)
  int64_t hb_blocks = cdiv64(TB, HB_ROWS);
  if (hb_blocks > 4 * (int64_t)ctx->sm_count) hb_blocks = 4 * (int64_t)ctx->sm_count;
  if (!rec_has_persistent_kernel(H) && g.C <= 4 && (K2 & 3) == 0 && hb_ncg <= 256 && hb_smem <= 200 * 1024 &&
      hb_blocks * (g.C * K2 + 4) <= gemm_splitk_ws_floats(H4, maxI, 2) &&
      ((reinterpret_cast<uintptr_t>(hL) | reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(part)) & 15) == 0) {
    if (!wgg_smem_ok(ctx, head_bwd_fused_kernel, 200 * 1024))  // once per context: the limit must cover every hidden size
      return wgg_fail(ctx, WGG_ECUDA, "head_bwd_fused_kernel: cannot reserve shared memory%s");
    {
      ProfScope prof(ctx, "head_bwd_fused_kernel", st, 4.0 * TB * K2 * g.C, 8.0 * TB * K2, "head_bwd_fused_kernel");
      head_bwd_fused_kernel<<<(unsigned)hb_blocks, 256, hb_smem, st>>>(out, dout, hL, params + g.off_wo, dh, part, g.T, B, K2, g.C);
      WGG_CHECK_LAUNCH(ctx, "head_bwd_fused_kernel");
    }
    head_bwd_reduce_kernel<<<(unsigned)cdiv64(g.C * K2 + 4, 256), 256, 0, st>>>(part, (int)hb_blocks, K2, g.C,
                                                                               dparams + g.off_wo, dparams + g.off_bo);
    WGG_CHECK_LAUNCH(ctx, "head_bwd_reduce_kernel");
  } else {
  head_bwd_kernel<<<ew_grid(TB * g.C), 256, 0, st>>>(out, dout, dpre, g.T, B, g.C);
  WGG_CHECK_LAUNCH(ctx, "head_bwd_kernel");
  {
    GemmP p;  // dWo (C x 2H) += dpre^T * hL
    p.tag = "gemm_kernel/head_wgrad";
    p.A = dpre; p.M = g.C; p.K = TB; p.sam = 1; p.sak = g.C;
    p.B = hL; p.N = 2 * H; p.sbk = 2 * H; p.sbn = 1;
    p.C = dparams + g.off_wo; p.scm = 2 * H; p.scn = 1; p.accumulate = 1; p.force_fp32 = 1;
    p.splitk = gemm_choose_splitk(ctx, p.M, p.N, p.K, 1); p.partial = part;
    WGG_TRY(gemm_launch(ctx, p, st));
    WGG_TRY(colsum_launch(ctx, dpre, TB, g.C, g.C, 1, 0, dparams + g.off_bo, nullptr, 0, 1, csws, st));
    GemmP q;  // dh (TB x 2H) = dpre * Wo
    q.tag = "gemm_kernel/head_dgrad";
    q.A = dpre; q.M = TB; q.K = g.C; q.sam = g.C; q.sak = 1;
    q.B = params + g.off_wo; q.N = 2 * H; q.sbk = 2 * H; q.sbn = 1;
    q.C = dh; q.scm = 2 * H; q.scn = 1; q.force_fp32 = 1;
    WGG_TRY(gemm_launch(ctx, q, st));
  }
  }
  for (int l = g.L - 1; l >= 0; --l) {
    const int I = g.in_dim(l);
    const float* lp = params + g.layer_off[l];
    float* dlp = dparams + g.layer_off[l];
    float* da = sv.gates[l];
    const float* in = l == 0 ? sv.x0 : sv.hseq[l - 1];
    // the forward's layout decision for this layer, re-derived (same context mode, same buffers)
    const bool ch = stash_chunk_enabled() && layer_chunked(ctx, g, B, l, in, lp, da, sv.hseq[l], zb_stash);
    if (ch) WGG_TRY(lstm_step_tc_backward(ctx, H, da, sv.cseq[l], lp, g.dir_stride[l], g.off_whh[l], dh, rec_scratch, g.T, B, 1, st));
    else WGG_TRY(rec_bwd_launch(ctx, H, da, sv.cseq[l], lp, g.dir_stride[l], g.off_whh[l], dh, rec_scratch, g.T, B, st));
    const bool tc_ih = wtc && I >= 128 && (I & 3) == 0;   // layer 0 (I0 = C + Z columns) stays on the mma.sync engine
    const bool tc_hh = wtc && H >= 128 && g.T > 1;
    // Layer 0 on the chunked stash: its input x0 = [prototype(t, b) | z(b)] makes dW_ih and dz cheap - the prototype columns are
    // pd weighted column sums of da (second planes of the unchunk pass), the latent columns and dz only need the time-summed da
    // (K = B instead of K = T B) - instead of two K = T B / N = 35 contractions on the mma.sync engine.
    const bool l0_fused = ch && l == 0 && g.pd <= 3 && g.T >= 4;
    if (ch) {
      // chunked da -> K-major image (weight gradients) + row-major copy (input gradient, bias sums, layer-0 weight gradient)
      dim3 ug((unsigned)(g.T * (pad128(B) / 128)), (unsigned)(H4 / 32), 2);
      ProfScope prof(ctx, "transpose_tf32_kernel", st, 0.0, 12.0 * (double)TB * H4 * 2, "unchunk_da_kernel");
      // (layer 0 with every consumer of the row-major copy gone - fused dW_ih / dz, tcgen05 dW_hh - does not write it)
      unchunk_da_kernel<<<ug, 256, 0, st>>>(da, daT, (l0_fused && tc_hh) ? nullptr : da_rm, colpart, g.T, B, H4,
                                            l0_fused ? sv.x0 : nullptr, g.I0, g.pd);
      WGG_CHECK_LAUNCH(ctx, "unchunk_da_kernel");
      if (l0_fused) {  // time-summed da (row-major [2][B][4H]) into the free dx buffer
        sum_t_chunk_kernel<<<dim3((unsigned)(pad128(B) / 128), (unsigned)(H4 / 4), 2), 128, 0, st>>>(da, dx, g.T, B, H4);
        WGG_CHECK_LAUNCH(ctx, "sum_t_chunk_kernel");
      }
      da = da_rm;
    }
    if (tc_ih || tc_hh) {
      if (!ch) WGG_TRY(transpose_tf32_launch(ctx, da, H4, TB * H4, daT, TB, TB * H4, TB, H4, 2, st));
      if (l == g.L - 1 && tc_hh) WGG_TRY(transpose_tf32_launch(ctx, sv.hseq[l], 2 * H, 0, hT[hcur], TB, 0, TB, 2 * H, 1, st));
      if (tc_ih) WGG_TRY(transpose_tf32_launch(ctx, in, I, 0, hT[hcur ^ 1], TB, 0, TB, I, 1, st));
    }
    if (tc_ih) {
      GemmP p;  // dW_ih[d] (4H x I) += da[d]^T (4H x TB, K-major image) * in^T (I x TB)^T
      p.tag = "gemm_tc/lstm_dWih";
      p.A = daT; p.M = H4; p.K = TB; p.sam = TB; p.sak = 1;
      p.B = hT[hcur ^ 1]; p.N = I; p.sbk = 1; p.sbn = TB;
      p.C = dlp; p.scm = I; p.scn = 1; p.accumulate = 1;
      p.nbatch = 2; p.bsA = TB * H4; p.bsB = 0; p.bsC = g.dir_stride[l];
      p.splitk = 2; p.partial = part;  // > 1: the tcgen05 engine chooses its own split count
      WGG_TRY(gemm_launch(ctx, p, st));
    } else if (l0_fused) {
      const int nblk = (int)(g.T * (pad128(B) / 128));
      for (int k = 0; k < g.pd; ++k) {  // dW_ih[d][:, k] += second stage over plane 1 + k
        colpart_reduce_kernel<<<dim3((unsigned)(H4 / 32), 2), 256, 0, st>>>(colpart + (int64_t)(1 + k) * 2 * nblk * H4, nblk, H4,
                                                                         dlp + k, nullptr, g.dir_stride[l], I);
        WGG_CHECK_LAUNCH(ctx, "colpart_reduce_kernel");
      }
      GemmP p;  // dW_ih[d][:, pd:] (4H x Z) += S[d]^T (4H x B) * z (B x Z);  z = the latent columns of x0 at t = 0
      p.tag = "gemm_kernel/lstm_dWih0_z";
      p.A = dx; p.M = H4; p.K = B; p.sam = 1; p.sak = H4;
      p.B = sv.x0 + g.pd; p.N = g.Z; p.sbk = I; p.sbn = 1;
      p.C = dlp + g.pd; p.scm = I; p.scn = 1; p.accumulate = 1; p.force_fp32 = 1;
      p.nbatch = 2; p.bsA = B * H4; p.bsB = 0; p.bsC = g.dir_stride[l];
      WGG_TRY(gemm_launch(ctx, p, st));
    } else {
      GemmP p;  // dW_ih[d] (4H x I) += da[d]^T * in
      p.tag = "gemm_kernel/lstm_dWih";
      p.A = da; p.M = H4; p.K = TB; p.sam = 1; p.sak = H4;
      p.B = in; p.N = I; p.sbk = I; p.sbn = 1;
      p.C = dlp; p.scm = I; p.scn = 1; p.accumulate = 1;
      p.nbatch = 2; p.bsA = TB * H4; p.bsB = 0; p.bsC = g.dir_stride[l];
      p.splitk = gemm_choose_splitk(ctx, p.M, p.N, p.K, 2); p.partial = part;
      WGG_TRY(gemm_launch(ctx, p, st));
    }
    if (tc_hh) {
      // dW_hh[d] (4H x H) += da[d][t]^T * h[d][t_prev]: in the transposed images the time shift is a column offset of B
      // (direction 0: da columns from B on against h columns from 0; direction 1 the other way round, h rows H..2H-1)
      GemmP p;
      p.tag = "gemm_tc/lstm_dWhh";
      p.A = daT + B; p.M = H4; p.K = (int64_t)(g.T - 1) * B; p.sam = TB; p.sak = 1;
      p.B = hT[hcur]; p.N = H; p.sbk = 1; p.sbn = TB;
      p.C = dlp + g.off_whh[l]; p.scm = H; p.scn = 1; p.accumulate = 1;
      p.nbatch = 2; p.bsA = TB * H4 - B; p.bsB = (int64_t)H * TB + B; p.bsC = g.dir_stride[l];
      p.splitk = 2; p.partial = part;
      WGG_TRY(gemm_launch(ctx, p, st));
    } else if (g.T > 1) {
      GemmP p;  // dW_hh[d] (4H x H) += da[d][t]^T * h[d][t_prev]; time shift = pointer offset
      p.tag = "gemm_kernel/lstm_dWhh";
      p.A = da + B * H4; p.M = H4; p.K = (int64_t)(g.T - 1) * B; p.sam = 1; p.sak = H4;
      p.B = sv.hseq[l]; p.N = H; p.sbk = 2 * H; p.sbn = 1;
      p.C = dlp + g.off_whh[l]; p.scm = H; p.scn = 1; p.accumulate = 1;
      p.nbatch = 2; p.bsA = TB * H4 - B * H4; p.bsB = B * 2 * H + H; p.bsC = g.dir_stride[l];
      p.splitk = gemm_choose_splitk(ctx, p.M, p.N, p.K, 2); p.partial = part;
      WGG_TRY(gemm_launch(ctx, p, st));
    }
    if (tc_ih) hcur ^= 1;  // the input's image is the next (lower) layer's output image
    // db_ih[d] = db_hh[d] += column sums of da[d] (chunked: second stage over the partials of the unchunk pass)
    if (ch) {
      colpart_reduce_kernel<<<dim3((unsigned)(H4 / 32), 2), 256, 0, st>>>(colpart, (int)(g.T * (pad128(B) / 128)), H4,
                                                                       dlp + g.off_bih[l], dlp + g.off_bhh[l], g.dir_stride[l], 1);
      WGG_CHECK_LAUNCH(ctx, "colpart_reduce_kernel");
    } else
      WGG_TRY(colsum_launch(ctx, da, TB, H4, H4, 2, TB * H4, dlp + g.off_bih[l], dlp + g.off_bhh[l], g.dir_stride[l], 1,
                            csws, st));
    if (l0_fused) {
      if (dz) {
        for (int d = 0; d < 2; ++d) {
          GemmP p;  // dz (B x Z) (+)= S[d] (B x 4H) * W_ih[d][:, pd:]
          p.tag = "gemm_kernel/lstm_dz";
          p.A = dx + (int64_t)d * B * H4; p.M = B; p.K = H4; p.sam = H4; p.sak = 1;
          p.B = lp + d * g.dir_stride[l] + g.pd; p.N = g.Z; p.sbk = I; p.sbn = 1;
          p.C = dz; p.scm = g.Z; p.scn = 1; p.accumulate = d; p.force_fp32 = 1;
          WGG_TRY(gemm_launch(ctx, p, st));
        }
      }
      return WGG_OK;  // layer 0 was the last layer of the loop; dz is complete
    }
    if (l > 0 || dz) {
      const bool tc_dx = wtc && I >= 128 && (I & 3) == 0;
      if (tc_dx) WGG_TRY(transpose_image_launch(ctx, lp, g.dir_stride[l], wihT, H4, I, 2, st));  // W_ih^T [d][I][4H]
      for (int d = 0; d < 2; ++d) {
        GemmP p;  // dx (TB x I) (+)= da[d] * W_ih[d]
        p.tag = tc_dx ? "gemm_tc/lstm_dx" : "gemm_kernel/lstm_dx";
        p.A = da + (int64_t)d * TB * H4; p.M = TB; p.K = H4; p.sam = H4; p.sak = 1;
        if (tc_dx) { p.B = wihT + (int64_t)d * I * H4; p.sbk = 1; p.sbn = H4; }
        else { p.B = lp + d * g.dir_stride[l]; p.sbk = I; p.sbn = 1; }
        p.N = I;
        p.C = dx; p.scm = I; p.scn = 1; p.accumulate = d;
        WGG_TRY(gemm_launch(ctx, p, st));
      }
      float* t = dh; dh = dx; dx = t;  // dh now holds d(input of layer l) = d(output of layer l-1)
    }
  }
  if (dz) {
    dz_kernel<<<(unsigned)cdiv64(B * g.Z, 256), 256, 0, st>>>(dh, dz, g.T, B, g.I0, g.pd, g.Z);
    WGG_CHECK_LAUNCH(ctx, "dz_kernel");
  }
  return WGG_OK;
}

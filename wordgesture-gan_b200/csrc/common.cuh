// Shared host/device helpers for libwgg_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <unordered_set>

#include "wgg.h"

struct wgg_ctx {
  int device = 0;
  int sm_count = 148;
  int math_mode = 0;
  int64_t launches = 0;
  // device scratch for reduction partials (allocated once in wgg_create): a ring of slots so that
  // consecutive reductions queued on one stream never share a slot with a finalize still in flight.
  float* red_scratch = nullptr;
  int red_slot[2] = {0, 0};
  int lane = 0;  // which of the two concurrently driven streams the next calls belong to (wgg_set_lane)
  // device word set by a persistent kernel whose pipeline wedged (bounded mbarrier waits); see wgg_async_error
  int* async_err = nullptr;
  // optional per-kernel-class CUDA-event timing (bench.py's roofline): see wgg_profile_enable
  bool prof_on = false;
  char prof_filter[64] = {0};
  static constexpr int kProfMax = 16384;
  cudaEvent_t* prof_ev = nullptr;  // 2*kProfMax events, created lazily
  int prof_n = 0;
  double prof_flops = 0.0, prof_bytes = 0.0;
  const char** prof_tag = nullptr;   // per event pair: static call-site tag
  double* prof_fl = nullptr;         // per event pair: FLOPs
  // kernels whose dynamic-shared-memory limit has been raised ON THIS DEVICE (function attributes are per device;
  // one ctx per device, see wgg_smem_ok)
  std::unordered_set<const void*> smem_cfg;
  char err[512] = {0};
};

// Raises a kernel's dynamic shared memory limit once per context (= per device).
template <typename F>
inline bool wgg_smem_ok(wgg_ctx* ctx, F* fn, size_t smem) {
  const void* key = reinterpret_cast<const void*>(fn);
  if (ctx->smem_cfg.count(key)) return true;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
  ctx->smem_cfg.insert(key);
  return true;
}

// RAII bracket: records a CUDA event pair around one launch of a profiled kernel class.
struct ProfScope {
  wgg_ctx* c;
  cudaStream_t st;
  int idx = -1;
  ProfScope(wgg_ctx* ctx, const char* name, cudaStream_t s, double flops, double bytes, const char* tag = nullptr)
      : c(ctx), st(s) {
    if (!c->prof_on || !strstr(name, c->prof_filter) || c->prof_n >= wgg_ctx::kProfMax) return;
    idx = c->prof_n++;
    c->prof_tag[idx] = tag ? tag : name;
    c->prof_fl[idx] = flops;
    c->prof_flops += flops;
    c->prof_bytes += bytes;
    cudaEventRecord(c->prof_ev[2 * idx], st);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(c->prof_ev[2 * idx + 1], st);
  }
};
constexpr int kRedBlocks = 296;  // 2 x 148 SMs
constexpr int kRedSlots = 64;
inline float* wgg_next_partial(wgg_ctx* ctx) {
  // each lane rotates through its own half of the ring
  const int half = kRedSlots / 2;
  float* p = ctx->red_scratch + (size_t)(ctx->lane * half + ctx->red_slot[ctx->lane]) * kRedBlocks;
  ctx->red_slot[ctx->lane] = (ctx->red_slot[ctx->lane] + 1) % half;
  return p;
}

inline int wgg_fail(wgg_ctx* ctx, int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  if (ctx) snprintf(ctx->err, sizeof(ctx->err), fmt, a, b, c);
  return code;
}

inline int wgg_check_launch(wgg_ctx* ctx, const char* name) {
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(ctx->err, sizeof(ctx->err), "%s: %s", name, cudaGetErrorString(e));
    return WGG_ECUDA;
  }
  return WGG_OK;
}

#define WGG_CHECK_LAUNCH(ctx, name)             \
  do {                                          \
    int rc__ = wgg_check_launch((ctx), (name)); \
    if (rc__ != WGG_OK) return rc__;            \
  } while (0)

#define WGG_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != WGG_OK) return rc__; \
  } while (0)

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr float kLeak = 0.2f;  // nn.LeakyReLU(0.2): src/gan/models.py:43,200,272

enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_TANH = 2 };

// ----------------------------------------------------------------------------------------------
// generic contraction engine (gemm.cu)
// C(m,n) = act( sum_k A(m,k) * B(k,n) + bias[n] + bias2[n] ),  fully strided operands, optional
// batching, optional deterministic split-K, optional conv1d "sliding window" view of one operand.
// ----------------------------------------------------------------------------------------------
struct GemmP {
  const float* A = nullptr;
  const float* B = nullptr;
  float* C = nullptr;
  int64_t M = 0;
  int64_t N = 0;
  int64_t K = 0;
  int64_t sam = 0, sak = 0;  // A(m,k) = A[m*sam + k*sak]
  int64_t sbk = 0, sbn = 0;  // B(k,n) = B[k*sbk + n*sbn]
  int64_t scm = 0, scn = 1;  // C(m,n) = C[m*scm + n*scn]
  int nbatch = 1;
  int64_t bsA = 0, bsB = 0, bsC = 0, bsBias = 0;
  const float* bias = nullptr;
  const float* bias2 = nullptr;
  int act = ACT_NONE;
  int accumulate = 0;  // C += result
  // conv window: operand X (A if conv_mode==1 with (row,col)=(m,k); B if conv_mode==2 with (row,col)=(k,n))
  // is a (rows = B*T, cols = ksize*Cin) sliding-window view of a channel-last (B,T,Cin) tensor:
  // X(r,c) = base[(r - pad)*Cin + c] if 0 <= (r % T) + c / Cin - pad < T else 0.
  int conv_mode = 0;
  int conv_T = 0, conv_Cin = 0, conv_pad = 0;
  // split-K: partial sums go to `partial` ([nbatch][splitk][M][N] dense) and are reduced deterministically.
  int splitk = 1;
  float* partial = nullptr;
  // numerics: 0 = follow the context's math mode (TF32 tensor cores when wgg_set_math_mode(ctx, 1));
  //           1 = always fp32 FMA (nn.Linear layers: the reference keeps cuBLAS matmuls in true fp32,
  //               only cuDNN LSTM / conv run TF32 - SURVEY.md 2.4 K1/K4/K7)
  int force_fp32 = 0;
  const char* tag = nullptr;  // call-site label for the profiler
  int x3 = 0;  // tensor-core mode: error-compensated 3xTF32 (always on for conv windows)
  // fp32 kernels only (nbatch == 1): rowsum[m] += sum_k A(m,k) in the same pass (bias gradient of a Linear layer's
  // weight-gradient GEMM); with split-K the partial buffer holds M extra floats per split.
  float* rowsum = nullptr;
  // tcgen05 engine only: write C (row m = t * chunk_B + b; bsC between batch entries) in the gate-buffer order of the
  // persistent H = 128 recurrence, [t][b / 128][N / 4][b % 128][4] - callers check gemm_tc_usable() first
  int out_chunk = 0;
  int64_t chunk_B = 0;
};

int gemm_launch(wgg_ctx* ctx, const GemmP& p, cudaStream_t st);
// out (+)= scale * sum(partial[0..n)) in fixed order (loss.cu); second stage of every deterministic loss reduction
int wgg_loss_finalize(wgg_ctx* ctx, const float* partial, int n, float scale, int accumulate, float* out, cudaStream_t st);
// tcgen05 / TMA engine for large "NT" contractions in the tensor-core math modes (gemm_tc.cu); gemm_launch routes to it
bool gemm_tc_usable(const wgg_ctx* ctx, const GemmP& p);
int gemm_tc_launch(wgg_ctx* ctx, const GemmP& p, cudaStream_t st);
// fused per-timestep kernels of the scaled recurrence (tcgen05 recurrent product + LSTM cell in the epilogue, gemm_tc.cu)
bool lstm_step_tc_usable(const wgg_ctx* ctx, int H, const float* gates, const float* hseq, const float* lp, int64_t off_whh,
                         int64_t dir_stride);
// chunked = 1: gate buffer / c in the chunked order [..][tile][columns / 4][128][4] (see gemm_tc.cu)
int lstm_step_tc_forward(wgg_ctx* ctx, int H, float* gates, const float* lp, int64_t dir_stride, int64_t off_whh, float* hseq,
                         float* cseq, float* cstate, int T, int64_t B, int store, int chunked, cudaStream_t st);
// persistent recurrence for gen_hidden_dim = 128 (one launch per layer; gemm_tc.cu)
bool lstm128_persist_usable(const wgg_ctx* ctx, int H, const float* gates, const float* hseq, const float* lp, int64_t off_whh,
                            int64_t dir_stride);
bool lstm128_persist_rowmajor();
// chunked = 1 (no-grad passes): gates is [2][T][ceil(B/128)][4H/4][128][4] as written by GemmP::out_chunk / xproj0_chunk
int lstm128_persist_forward(wgg_ctx* ctx, float* gates, const float* lp, int64_t dir_stride, int64_t off_whh, float* hseq,
                            float* cseq, int T, int64_t B, int store, int chunked, cudaStream_t st);
// chunked = 1: gates / cseq in the chunked stash order of the persistent H = 128 forward
int lstm_step_tc_backward(wgg_ctx* ctx, int H, float* gates, const float* cseq, const float* lp, int64_t dir_stride,
                          int64_t off_whh, const float* dh_out, float* scratch, int T, int64_t B, int chunked, cudaStream_t st);
// K-major (transposed, TF32-rounded) operand images + tcgen05 split-K GEMMs for the weight / input gradients of the scaled path
bool lstm_wgrad_tc_usable(const wgg_ctx* ctx, int H, int64_t B, int T);
int transpose_tf32_launch(wgg_ctx* ctx, const float* in, int64_t ld_in, int64_t bs_in, float* out, int64_t ld_out,
                          int64_t bs_out, int64_t R, int C, int nbatch, cudaStream_t st);
int transpose_image_launch(wgg_ctx* ctx, const float* in, int64_t in_bs, float* out, int N, int K, int nbatch, cudaStream_t st);
// workspace (floats) a split-K GEMM of this shape may need
int64_t gemm_splitk_ws_floats(int64_t M, int64_t N, int nbatch);
int gemm_choose_splitk(wgg_ctx* ctx, int64_t M, int64_t N, int64_t K, int nbatch);

// out[b][i] (+)= sum_s P[b][s][i]   (and the same into out2 if given)
int reduce_partials_launch(wgg_ctx* ctx, const float* P, int S, int64_t n, int nbatch, int64_t bsP, float* out,
                           float* out2, int64_t bsOut, int accumulate, cudaStream_t st);
// out[n] (+)= sum_m X[m*ldx + n]  (two-stage, deterministic); ws >= colsum_ws_floats(N)
int64_t colsum_ws_floats(int64_t N, int nbatch);
int colsum_launch(wgg_ctx* ctx, const float* X, int64_t M, int64_t N, int64_t ldx, int nbatch, int64_t bsX,
                  float* out, float* out2, int64_t bsOut, int accumulate, float* ws, cudaStream_t st);

// ----------------------------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in every thread. blockDim.x must be a multiple of 32 and <= 1024.
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 33 floats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? sh[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

__device__ __forceinline__ float leaky_f(float x) { return x > 0.f ? x : kLeak * x; }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// ----------------------------------------------------------------------------------------------
// shared elementwise launchers (ew.cu)
// ----------------------------------------------------------------------------------------------
// d[i] = (d[i] + (add ? add[i] : 0)) * (y[i] > 0 ? 1 : 0.2)      LeakyReLU backward (+ feature-grad injection)
int leaky_bwd_launch(wgg_ctx* ctx, const float* y, float* d, const float* add, int64_t n, cudaStream_t st);
int fill_launch(wgg_ctx* ctx, float* x, float v, int64_t n, cudaStream_t st);
inline int ew_blocks(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

// ----------------------------------------------------------------------------------------------
// dense-layer helpers (encoder.cu) shared with disc.cu
// ----------------------------------------------------------------------------------------------
int wgg_linear_fwd(wgg_ctx* ctx, const float* A, int64_t lda, const float* W, const float* bias, float* C,
                   int64_t ldc, int64_t M, int N, int K, int act, cudaStream_t st);
int wgg_linear_wgrad(wgg_ctx* ctx, const float* dY, int64_t ldy, const float* A, int64_t lda, float* dW, float* db, int64_t M,
                     int N, int K, int accumulate, float* part, cudaStream_t st);
int wgg_linear_dgrad(wgg_ctx* ctx, const float* dY, int64_t ldy, const float* W, float* dA, int64_t lda, int64_t M,
                     int N, int K, int accumulate, cudaStream_t st);

// ----------------------------------------------------------------------------------------------
// discriminator feature table (disc.cu) used by the feature-matching loss (loss.cu)
// ----------------------------------------------------------------------------------------------
struct FeatTable {
  int n = 0;
  int64_t off[WGG_MAX_HIDDEN_LAYERS + 2];    // float offset of the [B, width] block in the stash
  int64_t count[WGG_MAX_HIDDEN_LAYERS + 2];  // B * width
  int width[WGG_MAX_HIDDEN_LAYERS + 2];
};
int disc_feature_table(const wgg_model_cfg* cfg, int64_t B, FeatTable* ft);


// ----------------------------------------------------------------------------------------------
// tcgen05 persistent BiLSTM forward (lstm_tc.cu)
// ----------------------------------------------------------------------------------------------
bool generator_tc_supported(const wgg_model_cfg* cfg);
int64_t generator_tc_workspace_floats(const wgg_model_cfg* cfg, int64_t B);
int64_t generator_tc_bwd_workspace_floats(const wgg_model_cfg* cfg, int64_t B);
int64_t generator_tc_stash_floats(const wgg_model_cfg* cfg, int64_t B);
int generator_forward_tc(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const int64_t* layer_off,
                         const int64_t* dir_stride, const int64_t* off_whh, const int64_t* off_bih,
                         const int64_t* off_bhh, int64_t off_wo, int64_t off_bo, const float* proto, const float* z,
                         int64_t B, float* out, float* ws, int64_t ws_floats, float* stash, cudaStream_t st);
int generator_backward_tc_layers(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, float* dparams,
                                 const int64_t* layer_off, const int64_t* dir_stride, const int64_t* off_whh,
                                 const int64_t* off_bih, const int64_t* off_bhh, int64_t B, const float* stash,
                                 const float* out, const float* dout, int64_t off_wo, int64_t off_bo, float* dz,
                                 float* ws, int64_t ws_floats, cudaStream_t st);

// ----------------------------------------------------------------------------------------------
// tcgen05 conv1d layers of the TemporalDiscriminator (conv_tc.cu)
// ----------------------------------------------------------------------------------------------
// T = sequence length: a multiple of 128 (conv_tc_seq_ok); one MMA tile = 128 time steps of one gesture
bool conv_tc_seq_ok(int T);
int conv_tc_fwd_launch(wgg_ctx* ctx, const float* in, const float* wimg, const float* bias, float* out,
                       const float* act_lower, const float* dfeat, int64_t B, int T, int CinC, int taps, int pad, int N,
                       int mode, const char* tag, cudaStream_t st);
int64_t conv_tc_wgrad_ws_floats(wgg_ctx* ctx, int ncols_max);
int conv_tc_wgrad_launch(wgg_ctx* ctx, const float* dpre, const float* in, int64_t B, int T, int Cout, int Cin, int taps,
                         int pad, float* G, float* db, float* ws, cudaStream_t st);
int pack_x4_launch(wgg_ctx* ctx, const float* x, float* x4, int64_t B, int T, int C, cudaStream_t st);
int pool_fwd_chunk_launch(wgg_ctx* ctx, const float* a, float* pooled, int64_t B, int T, int C, cudaStream_t st);
int unpool_leaky_chunk_launch(wgg_ctx* ctx, const float* dpool, const float* a3, const float* dfeat, float* dpre,
                              int64_t B, int T, int C, cudaStream_t st);
int chunk_to_rows_launch(wgg_ctx* ctx, const float* in, float* out, int64_t B, int T, int C, cudaStream_t st);

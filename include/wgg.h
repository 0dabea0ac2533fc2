/*
 * wgg.h - C ABI of libwgg_sm100.so: the B200 (sm_100a) implementation of the WordGesture-GAN
 * training-step hot path.
 *
 * The reference (edwarddgao/WordGesture-GAN) is pure Python/PyTorch and has NO FFI of its own
 * (SURVEY.md 2.2); each entry point below therefore cites the reference *Python* interface whose
 * arithmetic it replaces (file:line relative to the reference root).  The Python host layer in
 * wordgesture-gan_b200/ binds these with ctypes from torch.autograd.Function.forward/backward
 * (see INTEGRATION.md for the reference-side stub).
 *
 * Conventions
 *   - plain pointers and sizes only; every tensor is fp32, contiguous, device memory owned by the caller;
 *   - every call is asynchronous on the cudaStream_t passed as `stream` (void*), never synchronises the
 *     host and never allocates device memory: scratch comes from the caller (`ws`, sized by the
 *     *_workspace_floats() queries) - so a whole training step is CUDA-graph capturable;
 *   - return value: 0 = WGG_OK, negative = error; wgg_last_error(ctx) gives the message;
 *   - one wgg_ctx per process/GPU; not thread-safe;
 *   - parameter vectors are FLAT: the module's tensors concatenated in `named_parameters()` order
 *     (the order torch.optim.Adam indexes them in, src/gan/trainer.py:60-79,208-211);
 *     spectral-norm buffers are flat in state_dict order (weight_u, weight_v per layer);
 *   - gradients w.r.t. parameters are ACCUMULATED (+=) into `dparams` (autograd `.grad` semantics).
 */
#ifndef WGG_H_
#define WGG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WGG_ABI_VERSION 1

#if defined(__GNUC__)
#define WGG_API __attribute__((visibility("default")))
#else
#define WGG_API
#endif

enum {
  WGG_OK = 0,
  WGG_EINVAL = -1,       /* bad argument / shape */
  WGG_ECUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
  WGG_EUNSUPPORTED = -3, /* configuration not covered by a compiled kernel */
  WGG_EWORKSPACE = -4    /* workspace too small */
};

#define WGG_MAX_HIDDEN_LAYERS 8

/* Mirror of ModelConfig (src/shared/config.py:11-33). */
typedef struct wgg_model_cfg {
  int32_t seq_length;         /* T, 128 */
  int32_t input_dim;          /* 3 */
  int32_t latent_dim;         /* Z, 32 */
  int32_t gen_hidden_dim;     /* H, 48 */
  int32_t gen_num_layers;     /* L, 4 */
  int32_t prototype_has_time; /* 0 -> generator sees (x,y) only */
  int32_t use_temporal_disc;  /* 1 -> TemporalDiscriminator (Conv1D), 0 -> MLP Discriminator */
  int32_t n_enc_hidden;
  int32_t enc_hidden_dims[WGG_MAX_HIDDEN_LAYERS];
  int32_t n_disc_hidden;
  int32_t disc_hidden_dims[WGG_MAX_HIDDEN_LAYERS];
} wgg_model_cfg;

typedef struct wgg_ctx wgg_ctx;

/* ---- context ------------------------------------------------------------------------------- */
WGG_API int wgg_abi_version(void);
WGG_API int wgg_create(wgg_ctx** out, int device);
WGG_API void wgg_destroy(wgg_ctx* ctx);
WGG_API const char* wgg_last_error(wgg_ctx* ctx);
/* number of kernels launched through this ctx since creation (bench.py's gpu_launches) */
WGG_API int64_t wgg_launch_count(wgg_ctx* ctx);
/* Per-kernel-class device timing for roofline reporting: every launch whose kernel name contains
 * `kernel_substr` ("gemm_kernel", "lstm_rec_fwd", ...) is bracketed by a CUDA event pair on its stream.
 * NULL disables.  wgg_profile_read synchronises on the recorded events and returns the summed duration,
 * the launch count and the algorithmic FLOPs / bytes those launches accounted for.  (max 16384 launches) */
WGG_API int wgg_profile_enable(wgg_ctx* ctx, const char* kernel_substr);
WGG_API int wgg_profile_read(wgg_ctx* ctx, double* total_ms, int64_t* launches, double* flops, double* bytes);
/* per call-site text report of the same brackets: one "tag launches ms gflop" line per tag */
WGG_API int wgg_profile_report(wgg_ctx* ctx, char* buf, int64_t size);
/* SYNCHRONISING debug query: reads (and clears) the device-side error word that the persistent tcgen05
 * kernels set when one of their bounded mbarrier waits times out (code = which wait).  0 = healthy. */
WGG_API int wgg_async_error(wgg_ctx* ctx, int* code);
/* math mode: 0 = fp32 FMA everywhere (default);
 * 1 = "tf32": LSTM and conv contractions on TF32 tensor cores (tcgen05 kernels; the numerics of the reference's
 *     own CUDA path, cuDNN TF32), nn.Linear layers stay fp32;
 * 2 = "tf32x3": LSTM as in 1, conv contractions in error-compensated 3xTF32 (fp32-grade gradients). */
WGG_API int wgg_set_math_mode(wgg_ctx* ctx, int mode);
/* A context may be driven from two CUDA streams at once (the two independent critic chains of the training step,
 * src/shared/utils.py:68-109, run concurrently).  Calls issued for the second stream are bracketed with
 * wgg_set_lane(ctx, 1) ... wgg_set_lane(ctx, 0) so that the library's internal reduction scratch of the two chains
 * never aliases (host-side state; no device work).  lane is 0 or 1. */
WGG_API int wgg_set_lane(wgg_ctx* ctx, int lane);

/* ---- Generator: replaces Generator.forward, src/gan/models.py:125-165 (+ its autograd) --------
 * params layout: nn.LSTM order - per layer, per direction: weight_ih (4H,I), weight_hh (4H,H),
 * bias_ih (4H), bias_hh (4H); then output_layer.weight (3,2H), output_layer.bias (3).
 * proto (B,T,3), z (B,Z) -> out (B,T,3).  stash == NULL: inference / no-grad (eval_gan.py:131-135). */
WGG_API int64_t wgg_generator_param_floats(const wgg_model_cfg* cfg);
WGG_API int64_t wgg_generator_stash_floats(const wgg_model_cfg* cfg, int64_t B);
WGG_API int64_t wgg_generator_workspace_floats(const wgg_model_cfg* cfg, int64_t B, int backward);
WGG_API int wgg_generator_forward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* proto,
                          const float* z, int64_t B, float* out, float* stash, float* ws, int64_t ws_floats,
                          void* stream);
/* dout (B,T,3) -> dparams (+=), dz (B,Z) (overwritten; may be NULL).  `stash` is consumed (overwritten). */
WGG_API int wgg_generator_backward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, int64_t B, float* stash,
                           const float* out, const float* dout, float* dparams, float* dz, float* ws,
                           int64_t ws_floats, void* stream);

/* ---- VariationalEncoder: replaces forward + reparameterize, src/gan/models.py:52-86 -----------
 * params: encoder.{0,2,..}.weight/bias, fc_mu.weight/bias, fc_log_var.weight/bias.
 * eps (B,Z) is the caller-drawn normal noise (torch.randn_like at models.py:85 - RNG stays in torch). */
WGG_API int64_t wgg_encoder_param_floats(const wgg_model_cfg* cfg);
WGG_API int64_t wgg_encoder_stash_floats(const wgg_model_cfg* cfg, int64_t B);
WGG_API int64_t wgg_encoder_workspace_floats(const wgg_model_cfg* cfg, int64_t B);
WGG_API int wgg_encoder_forward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* x,
                        const float* eps, int64_t B, float* z, float* mu, float* log_var, float* stash,
                        void* stream); /* stash (wgg_encoder_stash_floats) is required: it is also the activation scratch */
/* dz, dmu, dlog_var (B,Z) (any may be NULL = zero) -> dparams (+=), dx (B,T,3) (overwritten; may be NULL). */
WGG_API int wgg_encoder_backward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* x,
                         const float* eps, const float* log_var, int64_t B, const float* stash, const float* dz,
                         const float* dmu, const float* dlog_var, float* dparams, float* dx, float* ws,
                         int64_t ws_floats, void* stream);
/* The same pass with the KL term fused in (KLDivergenceLoss.forward, src/gan/losses.py:174-175, on the mu / log_var the
 * encoder just produced, trainer.py:161-171): both latent heads, the reparameterisation and the partial sums of
 * mean_b[-0.5 sum_j(1 + lv - mu^2 - exp(lv))] are ONE kernel; *kl (device scalar, may be NULL) receives the mean.
 * Backward: dkl (device scalar, may be NULL) is the upstream gradient of that scalar; mu is then required. */
WGG_API int wgg_encoder_forward_kl(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* x,
                           const float* eps, int64_t B, float* z, float* mu, float* log_var, float* stash, float* kl,
                           void* stream);
WGG_API int wgg_encoder_backward_kl(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* x,
                            const float* eps, const float* mu, const float* log_var, int64_t B, const float* stash,
                            const float* dz, const float* dmu, const float* dlog_var, const float* dkl, float* dparams,
                            float* dx, float* ws, int64_t ws_floats, void* stream);

/* ---- Discriminators: replace [Temporal]Discriminator.forward / get_all_features,
 * src/gan/models.py:202-243,293-353, including the spectral_norm pre-forward hook
 * (torch/nn/utils/spectral_norm.py:62-114).
 * params: per layer bias then weight_orig (named_parameters order).  uv: per layer weight_u, weight_v.
 * wgg_disc_spectral runs ONE power iteration per layer in place on uv when `training` (skipping the
 * output layer when with_output_layer == 0, as get_all_features does), and fills `sn` with the
 * effective weights W_orig/sigma (kernel-ready layouts) + the (u, v, sigma) snapshot backward needs. */
WGG_API int64_t wgg_disc_param_floats(const wgg_model_cfg* cfg);
WGG_API int64_t wgg_disc_uv_floats(const wgg_model_cfg* cfg);
WGG_API int64_t wgg_disc_sn_floats(const wgg_model_cfg* cfg);
WGG_API int64_t wgg_disc_stash_floats(const wgg_model_cfg* cfg, int64_t B);
WGG_API int64_t wgg_disc_workspace_floats(const wgg_model_cfg* cfg, int64_t B);
WGG_API int32_t wgg_disc_num_features(const wgg_model_cfg* cfg);
/* offset (floats) and per-sample width of feature k inside the stash (stash holds [B, width] blocks;
 * conv features are channel-last (B,T,C)). */
WGG_API int64_t wgg_disc_feature_offset(const wgg_model_cfg* cfg, int64_t B, int32_t k);
WGG_API int32_t wgg_disc_feature_width(const wgg_model_cfg* cfg, int32_t k);
WGG_API int wgg_disc_spectral(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, float* uv, int training,
                      int with_output_layer, float* sn, void* stream);
/* x (B,T,3) -> score (B,1) (NULL = features only), stash (features + pooled activations). */
WGG_API int wgg_disc_forward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* sn, const float* x,
                     int64_t B, float* score, float* stash, void* stream);
/* dscore (B,1) and/or dfeat (stash layout; only feature blocks are read) may be NULL.
 * dparams (+=) may be NULL (generator step: discriminator weight grads are discarded, utils.py:75,96);
 * dx (B,T,3) may be NULL (critic step). */
WGG_API int wgg_disc_backward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const float* sn, const float* x,
                      int64_t B, const float* stash, const float* dscore, const float* dfeat, float* dparams,
                      float* dx, float* ws, int64_t ws_floats, void* stream);
/* Conv feature block of a stash <-> the public (B, C*T) layout get_all_features returns (models.py:339).
 * The stash layout depends on the math mode the forward ran in ((B,T,C) or channel-chunked [B][C/4][T][4]);
 * do not change the math mode between a forward and the calls that consume its stash. */
WGG_API int wgg_disc_feature_convert(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* in, float* out, int64_t B,
                                     int32_t T, int32_t C, int to_stash, void* stream);
/* (B, T, C) -> (B, C, T) transpose (generic helper). */
WGG_API int wgg_transpose_tc(wgg_ctx* ctx, const float* in, float* out, int64_t B, int32_t T, int32_t C, void* stream);

/* ---- Losses: replace src/gan/losses.py ------------------------------------------------------
 * All scalar results are written to DEVICE floats (no host sync); backward entry points take the
 * upstream scalar gradient as a device float pointer `g` (NULL = 1.0). */
/* out[0] = scale * mean(x[0..n)) (+ out[0] if accumulate)            losses.py:43,58 */
WGG_API int wgg_mean(wgg_ctx* ctx, const float* x, int64_t n, float scale, int accumulate, float* out, void* stream);
/* dx[i] = g * scale / n                                                 */
WGG_API int wgg_mean_backward(wgg_ctx* ctx, const float* g, float scale, int64_t n, float* dx, void* stream);
/* out[0] = scale * mean|a-b|                                          losses.py:120,147 */
WGG_API int wgg_l1_mean(wgg_ctx* ctx, const float* a, const float* b, int64_t n, float scale, int accumulate, float* out,
                void* stream);
/* da[i] (+)= g * scale * sign(a-b) / n                                  */
WGG_API int wgg_l1_mean_backward(wgg_ctx* ctx, const float* a, const float* b, const float* g, float scale, int64_t n,
                         int accumulate, float* da, void* stream);
/* FeatureMatchingLoss over discriminator stashes (layout-invariant)    losses.py:86-93 */
WGG_API int wgg_feature_matching(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* real_stash, const float* fake_stash,
                         int64_t B, float scale, int accumulate, float* out, void* stream);
WGG_API int wgg_feature_matching_backward(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* real_stash,
                                  const float* fake_stash, const float* g, float scale, int64_t B, float* dfeat,
                                  void* stream);
/* KLDivergenceLoss                                                     losses.py:174-175 */
WGG_API int wgg_kl(wgg_ctx* ctx, const float* mu, const float* log_var, int64_t B, int32_t Z, float scale, int accumulate,
           float* out, void* stream);
WGG_API int wgg_kl_backward(wgg_ctx* ctx, const float* mu, const float* log_var, const float* g, float scale, int64_t B,
                    int32_t Z, float* dmu, float* dlog_var, void* stream);

/* ---- Optimiser step: replaces clip_grad_norm_ + Adam.step, src/shared/utils.py:87-88,108-109,132-135
 * One fused pass over the module's flat buffers: total L2 norm -> clip coefficient
 * min(1, max_norm/(norm+1e-6)) -> bias-corrected Adam (torch.optim.Adam, no weight decay/amsgrad).
 * `step` is the 1-based step count AFTER increment.  max_norm <= 0 disables clipping.
 * grad_norm_out: optional device float receiving the pre-clip norm.  g is scaled in place like
 * clip_grad_norm_ does.  ws needs wgg_clip_adam_workspace_floats() floats. */
WGG_API int64_t wgg_clip_adam_workspace_floats(void);
WGG_API int wgg_clip_adam(wgg_ctx* ctx, float* p, float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, int64_t step, float max_norm, float* grad_norm_out, float* ws,
                  void* stream);

/* Same update with the step counter (int32, incremented by the call) and the learning rate held in DEVICE memory,
 * so that nothing step-dependent is baked into the launch: this is the variant a captured CUDA graph replays. */
WGG_API int wgg_clip_adam_dev(wgg_ctx* ctx, float* p, float* g, float* m, float* v, int64_t n, const float* lr_dev,
                      float beta1, float beta2, float eps, int* step_dev, float max_norm, float* grad_norm_out,
                      float* ws, void* stream);

/* ---- generic building block exposed for tests/bench: C[M,N] = act(A[M,K] * W[N,K]^T + bias) ---- */
WGG_API int wgg_linear(wgg_ctx* ctx, const float* A, const float* W, const float* bias, float* C, int64_t M, int32_t N,
               int32_t K, int act /*0 none, 1 leaky(0.2), 2 tanh*/, void* stream);

/* ---- Data-parallel gradient exchange over peer memory (no counterpart in the single-GPU reference) ----
 * One-shot mean all-reduce of a flat gradient bucket: every rank reads all peers' buckets over NVLink and sums them in
 * rank order.  peer_grads / peer_flags: DEVICE arrays of `world` pointers - entry r is rank r's bucket (n floats) and
 * flag block (wgg_p2p_flag_words() uint32, zero-initialised) as mapped into THIS process (CUDA IPC; entry `rank` is
 * the local memory).  avg: n floats of local scratch; local_grad: the local bucket (receives the mean);
 * state: 4 zero-initialised uint32 of local device memory private to this bucket.  n % 4 == 0, world <= 16.
 * Every rank must issue the same sequence of calls per bucket; asynchronous on `stream`, CUDA-graph capturable
 * (the epoch lives in `state`); a peer that never arrives ends the wait after ~2 s and sets wgg_async_error. */
WGG_API int64_t wgg_p2p_flag_words(void);
/* Shared bucket memory: wgg_p2p_alloc = cudaMalloc (zero-filled) on ctx's device + its 64-byte CUDA IPC handle (send it
 * to the peers by any means); wgg_p2p_open maps a peer's handle into THIS rank's device address space (NVLink peer
 * access enabled on demand); wgg_p2p_close(ptr, imported) unmaps an imported bucket / frees an own one. */
WGG_API int wgg_p2p_alloc(wgg_ctx* ctx, int64_t bytes, void** ptr, unsigned char* handle64);
WGG_API int wgg_p2p_open(wgg_ctx* ctx, const unsigned char* handle64, void** ptr);
WGG_API int wgg_p2p_close(wgg_ctx* ctx, void* ptr, int imported);
WGG_API int wgg_p2p_allreduce_avg(wgg_ctx* ctx, const float* const* peer_grads, uint32_t* const* peer_flags, int rank, int world,
                                  int64_t n, float* avg, float* local_grad, uint32_t* state, void* stream);

/* ---- Evaluation metrics on the GPU: the kernels behind evaluate_all_metrics,
 * src/gan/evaluation.py:297-500.  n is the number of gestures; all arrays fp32 device memory. --------------------- */
/* out (na, nb): Euclidean distance matrix, replaces scipy cdist(a, b, 'euclidean') at evaluation.py:335,474-476;
 * a (na, d), b (nb, d) row-major (d = 2 T for the flattened (x, y) trajectories). */
WGG_API int wgg_eval_cdist(wgg_ctx* ctx, const float* a, int64_t na, const float* b, int64_t nb, int32_t d, float* out,
                           void* stream);
/* out[r] = np.sort(m, axis=1)[r, k] (the k-NN radius of evaluation.py:475,478); m (rows, cols); 0 <= k < 8. */
WGG_API int wgg_eval_row_kth(wgg_ctx* ctx, const float* m, int64_t rows, int64_t cols, int32_t k, float* out, void* stream);
/* out2 = {precision, recall} of evaluation.py:480-484 from rf = cdist(real, fake) (n_real, n_fake) and the two radius
 * vectors; ws2: 2 floats of scratch. */
WGG_API int wgg_eval_precision_recall(wgg_ctx* ctx, const float* rf, int64_t n_real, int64_t n_fake, const float* real_radii,
                                      const float* fake_radii, float* out2, float* ws2, void* stream);
/* out[0] = mean over gestures of mean_t sqrt((S x)_t^2 + (S y)_t^2) (evaluation.py:364-374); g (n, T, C >= 2);
 * S (T, T): the linear operator of savgol_filter(window, poly, deriv=3) built by the host; ws: n floats. */
WGG_API int wgg_eval_jerk(wgg_ctx* ctx, const float* g, int64_t n, int32_t T, int32_t C, const float* S, float* out, float* ws,
                          void* stream);
/* out4 = {velocity_corr, acceleration_corr, speed_profile_corr, time_delta_corr} of evaluation.py:162-305 for the
 * pairs (real_i, fake_i); real, fake (n, T, C >= 3) with time in channel 2; T <= 257; ws: 8 n floats. */
WGG_API int wgg_eval_dynamics(wgg_ctx* ctx, const float* real, const float* fake, int64_t n, int32_t T, int32_t C, float* out4,
                              float* ws, void* stream);

/* ---- Word prototypes / minimum-jerk trajectories on the GPU ---------------------------------
 * keys (n, maxk, 2) float64 key centres of each word (QWERTYKeyboard._get_key_positions, src/shared/keyboard.py:679-686),
 * nkeys (n) int32 -> out (n, T, 3) fp32 rows (x, y, t).  One thread block per word, float64 arithmetic like numpy. */
/* QWERTYKeyboard.get_word_prototype, keyboard.py:710-765: straight segments between key centres sampled at uniform arc
 * length, t = linspace(0, 1); maxk <= 64. */
WGG_API int wgg_word_prototypes(wgg_ctx* ctx, const double* keys, const int32_t* nkeys, int64_t n, int32_t maxk, int32_t T,
                                float* out, void* stream);
/* generate_minimum_jerk_trajectory, keyboard.py:389-514 (C2 quintic-Hermite path through the via-points, 1000-point fine
 * trajectory, arc-length resampling, time from the inverted s(tau)).  The reference's random draws are inputs:
 * key_noise (n, maxk, 2) = N(0, offset_std) for the interior keys (row i - 1 for key i, :426-429), mid_noise (n, maxk) =
 * N(0, offset_std / 2) per segment (:440-442); both NULL for offset_std = 0.  maxk <= 32. */
WGG_API int wgg_minimum_jerk(wgg_ctx* ctx, const double* keys, const int32_t* nkeys, int64_t n, int32_t maxk,
                             const double* key_noise, const double* mid_noise, int include_midpoints, int32_t T, float* out,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WGG_H_ */

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


// ---------------------------------------------------------------------------------------------
// One-shot mean all-reduce of a flat gradient bucket over peer memory (NVLink 5 / NVSwitch), for the data-parallel
// step: the buckets are small (276 KB per critic, 1.2 MB for G || E), so the collective is pure
// latency - every rank simply READS all W peers' buckets over NVLink and sums them in rank order (the same order on
// every rank: replicas stay bit-identical), instead of going through NCCL's ring / tree protocol.
// Each rank's bucket and flag block live in memory the peers have mapped (CUDA IPC); flags[0][r] / flags[1][r] are
// written by rank r (entry: "my gradients are complete", exit: "I have finished reading yours").  The launch:
//   entry barrier (system-scope release / acquire flags) -> avg = (1/W) sum_r grad_r into a LOCAL buffer ->
//   local grid barrier -> exit barrier -> local gradient bucket := avg   (what clip + Adam then consume).
// The epoch lives in device memory, so the launch is CUDA-graph capturable; every spin is bounded (~2 s) and reports
// through the context's async error word instead of hanging the GPU.  grid <= 32 CTAs: all co-resident.
// ---------------------------------------------------------------------------------------------
constexpr int kP2PMaxWorld = 16;
constexpr int kP2PBlocks = 32;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// spins until *p >= target (acquire at the given scope); false (and the error word set) after ~2 s
template <bool SYS>
__device__ bool spin_until(const uint32_t* p, uint32_t target, int* gerr, int code) {
  const long long t0 = clock64();
  while (true) {
    const uint32_t v = SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p);
    if ((int32_t)(v - target) >= 0) return true;
    if (clock64() - t0 > 4000000000ll) {
      atomicCAS(gerr, 0, code);
      return false;
    }
  }
}

// state[0] = epoch of the last completed launch, state[1] = monotonic CTA counter, state[2] / state[3] = gates
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const float* const* __restrict__ peer_grads,
                                                            uint32_t* const* __restrict__ peer_flags, int rank, int world,
                                                            int64_t n4, float* __restrict__ avg, float* __restrict__ local_grad,
                                                            uint32_t* __restrict__ state, int* __restrict__ gerr) {
  const int tid = threadIdx.x;
  const uint32_t e = state[0] + 1u;
  uint32_t* my_flags = peer_flags[rank];
  // ---- entry barrier: every rank's bucket is complete ----
  if (blockIdx.x == 0) {
    if (tid < world) {
      __threadfence_system();
      st_release_sys(peer_flags[tid] + rank, e);
      spin_until<true>(my_flags + tid, e, gerr, 61);
    }
    __syncthreads();
    if (tid == 0) st_release_gpu(state + 2, e);
  } else {
    if (tid == 0) spin_until<false>(state + 2, e, gerr, 62);
    __syncthreads();
  }
  // ---- avg = mean over ranks, summed in rank order (identical on every rank) ----
  const float inv = 1.f / (float)world;
  float4* avg4 = reinterpret_cast<float4*>(avg);
  for (int64_t i = (int64_t)blockIdx.x * 256 + tid; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 a = ld_sys_f4(reinterpret_cast<const float4*>(peer_grads[0]) + i);
    for (int r = 1; r < world; ++r) {
      const float4 b = ld_sys_f4(reinterpret_cast<const float4*>(peer_grads[r]) + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    avg4[i] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  }
  // ---- local grid barrier, then exit barrier: nobody overwrites a bucket a peer may still be reading ----
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicAdd(state + 1, 1u);
  }
  if (blockIdx.x == 0) {
    if (tid == 0) spin_until<false>(state + 1, e * gridDim.x, gerr, 63);
    __syncthreads();
    if (tid < world) {
      st_release_sys(peer_flags[tid] + kP2PMaxWorld + rank, e);
      spin_until<true>(my_flags + kP2PMaxWorld + tid, e, gerr, 64);
    }
    __syncthreads();
    if (tid == 0) st_release_gpu(state + 3, e);
  } else {
    if (tid == 0) spin_until<false>(state + 3, e, gerr, 65);
    __syncthreads();
  }
  // ---- the local bucket becomes the averaged gradient ----
  float4* g4 = reinterpret_cast<float4*>(local_grad);
  for (int64_t i = (int64_t)blockIdx.x * 256 + tid; i < n4; i += (int64_t)gridDim.x * 256) g4[i] = avg4[i];
  if (blockIdx.x == 0 && tid == 0) state[0] = e;  // every CTA has read the old epoch: they all passed the grid barrier
}

}  // namespace

extern "C" int64_t wgg_p2p_flag_words(void) { return 2 * kP2PMaxWorld; }

// Shared buckets are plain cudaMalloc allocations of the library (an IPC handle names a whole allocation, and the
// importing side must map it with ITS OWN device current so that the mapping lands in the address space its kernels
// run in; cudaIpcMemLazyEnablePeerAccess then enables NVLink peer access to the owner's device).
extern "C" int wgg_p2p_alloc(wgg_ctx* ctx, int64_t bytes, void** ptr, unsigned char* handle64) {
  if (!ctx || !ptr || !handle64 || bytes <= 0) return WGG_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return wgg_fail(ctx, WGG_ECUDA, "wgg_p2p_alloc: cudaSetDevice%s");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    return wgg_fail(ctx, WGG_ECUDA, "wgg_p2p_alloc: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return WGG_OK;
}

extern "C" int wgg_p2p_open(wgg_ctx* ctx, const unsigned char* handle64, void** ptr) {
  if (!ctx || !ptr || !handle64) return WGG_EINVAL;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return wgg_fail(ctx, WGG_ECUDA, "wgg_p2p_open: cudaSetDevice%s");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return wgg_fail(ctx, WGG_ECUDA, "wgg_p2p_open: %s", cudaGetErrorString(e));
  *ptr = p;
  return WGG_OK;
}

extern "C" int wgg_p2p_close(wgg_ctx* ctx, void* ptr, int imported) {
  if (!ctx || !ptr) return WGG_EINVAL;
  cudaSetDevice(ctx->device);
  const cudaError_t e = imported ? cudaIpcCloseMemHandle(ptr) : cudaFree(ptr);
  return e == cudaSuccess ? WGG_OK : wgg_fail(ctx, WGG_ECUDA, "wgg_p2p_close: %s", cudaGetErrorString(e));
}

extern "C" int wgg_p2p_allreduce_avg(wgg_ctx* ctx, const float* const* peer_grads, uint32_t* const* peer_flags, int rank,
                                     int world, int64_t n, float* avg, float* local_grad, uint32_t* state, void* stream) {
  if (!ctx || !peer_grads || !peer_flags || !avg || !local_grad || !state || world < 1 || world > kP2PMaxWorld || rank < 0 ||
      rank >= world || n <= 0 || (n & 3))
    return wgg_fail(ctx, WGG_EINVAL, "wgg_p2p_allreduce_avg: bad argument (n must be a multiple of 4, world <= 16)%s");
  const int64_t n4 = n / 4;
  // the grid size is part of the barrier arithmetic (state[1] counts CTAs): a function of n only
  int grid = (int)cdiv64(n4, 256 * 4);
  if (grid > kP2PBlocks) grid = kP2PBlocks;
  if (grid < 1) grid = 1;
  p2p_allreduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(peer_grads, peer_flags, rank, world, n4, avg, local_grad, state,
                                                              ctx->async_err);
  WGG_CHECK_LAUNCH(ctx, "p2p_allreduce_kernel");
  return WGG_OK;
}

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

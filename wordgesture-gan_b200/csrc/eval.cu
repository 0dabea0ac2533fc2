// Evaluation metrics of generated gestures on the GPU: the O(n^2 d) distance matrices, the k-NN
// manifold precision / recall, the Savitzky-Golay jerk and the four time-aware dynamics correlations of
// src/gan/evaluation.py:297-500.  All HBM-bound / latency-bound fp32 work with fp64 accumulation where the reference
// (numpy) accumulates in float64; results are deterministic (fixed reduction order, integer-valued atomics only).
#include <math.h>

#include "common.cuh"

namespace ev {

// ---------------------------------------------------------------------------------------------
// out[i][j] = || a_i - b_j ||_2      a (na, d), b (nb, d) row-major     (cdist(., ., 'euclidean'): evaluation.py:335,474-476)
// 32 x 32 output tile per 256-thread block (2 x 2 per thread), the d axis in slabs of 32 through shared memory;
// differences are formed before squaring, as scipy does (no |a|^2 + |b|^2 - 2ab cancellation).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cdist_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb,
                                                    int d, float* __restrict__ out) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < d; k0 += 32) {
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int r = e >> 5, c = e & 31;
      sa[r][c] = (i0 + r < na && k0 + c < d) ? __ldg(a + (int64_t)(i0 + r) * d + k0 + c) : 0.f;
      sb[r][c] = (j0 + r < nb && k0 + c < d) ? __ldg(b + (int64_t)(j0 + r) * d + k0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float a0 = sa[ty][k], a1 = sa[ty + 16][k], b0 = sb[tx][k], b1 = sb[tx + 16][k];
      float t;
      t = a0 - b0; acc[0][0] = fmaf(t, t, acc[0][0]);
      t = a0 - b1; acc[0][1] = fmaf(t, t, acc[0][1]);
      t = a1 - b0; acc[1][0] = fmaf(t, t, acc[1][0]);
      t = a1 - b1; acc[1][1] = fmaf(t, t, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = i0 + ty + 16 * p, j = j0 + tx + 16 * q;
      if (i < na && j < nb) out[(int64_t)i * nb + j] = sqrtf(acc[p][q]);
    }
}

// out[r] = the (k+1)-th smallest entry of row r  (np.sort(m, axis=1)[:, k]: evaluation.py:475,478), k < 8
constexpr int KTH_MAX = 8;
__global__ void __launch_bounds__(128) row_kth_kernel(const float* __restrict__ m, int cols, int k, float* __restrict__ out) {
  __shared__ float s[128 * KTH_MAX];
  const float* row = m + (int64_t)blockIdx.x * cols;
  float best[KTH_MAX];
#pragma unroll
  for (int i = 0; i < KTH_MAX; ++i) best[i] = INFINITY;
  for (int c = threadIdx.x; c < cols; c += 128) {
    float v = __ldg(row + c);
#pragma unroll
    for (int i = 0; i < KTH_MAX; ++i)  // insertion into the sorted list of the smallest KTH_MAX values seen
      if (v < best[i]) { const float t = best[i]; best[i] = v; v = t; }
  }
#pragma unroll
  for (int i = 0; i < KTH_MAX; ++i) s[threadIdx.x * KTH_MAX + i] = best[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    float sel[KTH_MAX];
#pragma unroll
    for (int i = 0; i < KTH_MAX; ++i) sel[i] = INFINITY;
    for (int e = 0; e < 128 * KTH_MAX; ++e) {
      float v = s[e];
#pragma unroll
      for (int i = 0; i < KTH_MAX; ++i)
        if (v < sel[i]) { const float t = sel[i]; sel[i] = v; v = t; }
    }
    float r = sel[0];
#pragma unroll
    for (int i = 1; i < KTH_MAX; ++i)
      if (i == k) r = sel[i];
    out[blockIdx.x] = r;
  }
}

// precision = mean_j any_i (rf[i][j] <= real_radii[i]);  recall = mean_i any_j (rf[i][j] <= fake_radii[j])
// (evaluation.py:480-484).  counts[0] / counts[1] receive integer-valued float sums (exact and order-independent).
__global__ void __launch_bounds__(256) prec_kernel(const float* __restrict__ rf, int nr, int nf, const float* __restrict__ real_radii,
                                                   float* __restrict__ counts) {
  // block = 256 consecutive fake columns j; loop over real rows (coalesced along j)
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= nf) return;
  bool hit = false;
  for (int i = 0; i < nr && !hit; ++i) hit = __ldg(rf + (int64_t)i * nf + j) <= __ldg(real_radii + i);
  if (hit) atomicAdd(counts, 1.f);
}
__global__ void __launch_bounds__(128) rec_kernel(const float* __restrict__ rf, int nf, const float* __restrict__ fake_radii,
                                                  float* __restrict__ counts) {
  __shared__ int s_hit;
  if (threadIdx.x == 0) s_hit = 0;
  __syncthreads();
  const float* row = rf + (int64_t)blockIdx.x * nf;
  bool hit = false;
  for (int j = threadIdx.x; j < nf; j += 128) hit |= __ldg(row + j) <= __ldg(fake_radii + j);
  if (hit) s_hit = 1;
  __syncthreads();
  if (threadIdx.x == 0 && s_hit) atomicAdd(counts + 1, 1.f);
}
__global__ void prec_rec_finalize_kernel(const float* __restrict__ counts, int nr, int nf, float* __restrict__ out) {
  out[0] = counts[0] / (float)nf;
  out[1] = counts[1] / (float)nr;
}

// ---------------------------------------------------------------------------------------------
// Savitzky-Golay jerk (evaluation.py:364-374): jerk[i] = mean_t sqrt((S x_i)_t^2 + (S y_i)_t^2), S = the (T, T) operator
// of savgol_filter(window, poly, deriv=3, mode='interp') built by the host.  One block per gesture.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) jerk_kernel(const float* __restrict__ g, int T, int C, const float* __restrict__ S,
                                                   float* __restrict__ out) {
  extern __shared__ float sm[];  // x[T], y[T]
  float* sx = sm;
  float* sy = sm + T;
  const float* p = g + (int64_t)blockIdx.x * T * C;
  for (int t = threadIdx.x; t < T; t += 128) {
    sx[t] = __ldg(p + t * C);
    sy[t] = __ldg(p + t * C + 1);
  }
  __syncthreads();
  double acc = 0.0;
  for (int t = threadIdx.x; t < T; t += 128) {
    const float* row = S + (int64_t)t * T;
    double dx = 0.0, dy = 0.0;
    for (int j = 0; j < T; ++j) {
      const double w = (double)__ldg(row + j);
      dx += w * (double)sx[j];
      dy += w * (double)sy[j];
    }
    acc += sqrt(dx * dx + dy * dy);
  }
  // fixed-order block sum in double
  __shared__ double sd[128];
  sd[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 64; w > 0; w >>= 1) {
    if (threadIdx.x < w) sd[threadIdx.x] += sd[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(sd[0] / T);
}

// ---------------------------------------------------------------------------------------------
// Time-aware dynamics correlations (evaluation.py:55-305).  One 256-thread block per (real_i, fake_i) pair, four
// metrics: 0 velocity (2(T-1) values, clipped to [p1, p99]), 1 acceleration (2(T-2), [p1, p99]), 2 speed profile
// (T-1, [0, p99]), 3 time deltas (T-1, unclipped).  Percentiles by a bitonic sort of the row (np.percentile's linear
// interpolation), Pearson correlation with fp64 sums.  val[i][m] = correlation, ok[i][m] = 1 if the reference would
// have kept it (both rows have std > 1e-10 and the correlation is not NaN).
// ---------------------------------------------------------------------------------------------
constexpr int DYN_MAX = 512;  // 2 (T - 1) <= 512  ->  T <= 257

__device__ __forceinline__ float safe_dt(float dt) {
  // np.where(|dt| > 1e-10, dt, 1e-10 * sign(dt + 1e-20))
  if (fabsf(dt) > 1e-10f) return dt;
  const float s = dt + 1e-20f;
  return 1e-10f * (s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f));
}

__device__ void bitonic_sort(float* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
}

__device__ double block_sum_d(double v, double* sh) {
  __syncthreads();
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int w = blockDim.x >> 1; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

__device__ float percentile_sorted(const float* sorted, int L, double q) {
  const double pos = (double)(L - 1) * q;
  const int lo = (int)floor(pos);
  const int hi = lo + 1 < L ? lo + 1 : lo;
  const double fr = pos - lo;
  // numpy's _lerp: a + (b - a) * t, switched to b - (b - a) * (1 - t) for t >= 0.5
  const double a = sorted[lo], b = sorted[hi];
  const double d = b - a;
  return (float)(fr >= 0.5 ? b - d * (1.0 - fr) : a + d * fr);
}

__global__ void __launch_bounds__(256) dynamics_kernel(const float* __restrict__ real, const float* __restrict__ fake, int T,
                                                       int C, float* __restrict__ val, float* __restrict__ ok) {
  __shared__ float s_t[2][DYN_MAX / 2 + 2];         // times of the two gestures
  __shared__ float s_xy[2][2][DYN_MAX / 2 + 2];     // [gesture][x|y][t]
  __shared__ float s_v[2][DYN_MAX];                 // current metric's rows (real, fake)
  __shared__ float s_sort[DYN_MAX];
  __shared__ double s_red[256];
  __shared__ float s_lim[2][2];
  const float* src[2] = {real + (int64_t)blockIdx.x * T * C, fake + (int64_t)blockIdx.x * T * C};
  for (int g = 0; g < 2; ++g)
    for (int t = threadIdx.x; t < T; t += 256) {
      s_xy[g][0][t] = __ldg(src[g] + t * C);
      s_xy[g][1][t] = __ldg(src[g] + t * C + 1);
      s_t[g][t] = __ldg(src[g] + t * C + 2);
    }
  __syncthreads();
  for (int metric = 0; metric < 4; ++metric) {
    const int L = metric == 0 ? 2 * (T - 1) : (metric == 1 ? 2 * (T - 2) : T - 1);
    // ---- build the two rows ----
    for (int g = 0; g < 2; ++g)
      for (int e = threadIdx.x; e < L; e += 256) {
        float r;
        if (metric == 0) {  // velocity, flattened (t, xy)
          const int t = e >> 1, c = e & 1;
          r = (s_xy[g][c][t + 1] - s_xy[g][c][t]) / safe_dt(s_t[g][t + 1] - s_t[g][t]);
        } else if (metric == 1) {  // acceleration
          const int t = e >> 1, c = e & 1;
          const float v0 = (s_xy[g][c][t + 1] - s_xy[g][c][t]) / safe_dt(s_t[g][t + 1] - s_t[g][t]);
          const float v1 = (s_xy[g][c][t + 2] - s_xy[g][c][t + 1]) / safe_dt(s_t[g][t + 2] - s_t[g][t + 1]);
          const float m0 = (s_t[g][t] + s_t[g][t + 1]) * 0.5f, m1 = (s_t[g][t + 1] + s_t[g][t + 2]) * 0.5f;
          r = (v1 - v0) / safe_dt(m1 - m0);
        } else if (metric == 2) {  // speed
          const float dtv = safe_dt(s_t[g][e + 1] - s_t[g][e]);
          const float vx = (s_xy[g][0][e + 1] - s_xy[g][0][e]) / dtv, vy = (s_xy[g][1][e + 1] - s_xy[g][1][e]) / dtv;
          r = sqrtf(vx * vx + vy * vy);
        } else {
          r = s_t[g][e + 1] - s_t[g][e];
        }
        s_v[g][e] = r;
      }
    __syncthreads();
    // ---- validity: np.std (population) of the raw rows ----
    bool valid = L > 1;
    for (int g = 0; g < 2; ++g) {
      double s1 = 0.0;
      for (int e = threadIdx.x; e < L; e += 256) s1 += (double)s_v[g][e];
      const double mean = block_sum_d(s1, s_red) / L;
      double s2 = 0.0;
      for (int e = threadIdx.x; e < L; e += 256) { const double dlt = (double)s_v[g][e] - mean; s2 += dlt * dlt; }
      const double var = block_sum_d(s2, s_red) / L;
      if (!(sqrt(var) > 1e-10)) valid = false;
    }
    // ---- clipping limits from the percentiles of each row ----
    if (metric < 3) {
      int n2 = 1;
      while (n2 < L) n2 <<= 1;
      for (int g = 0; g < 2; ++g) {
        for (int e = threadIdx.x; e < n2; e += 256) s_sort[e] = e < L ? s_v[g][e] : INFINITY;
        __syncthreads();
        bitonic_sort(s_sort, n2);
        if (threadIdx.x == 0) {
          s_lim[g][0] = metric == 2 ? 0.f : percentile_sorted(s_sort, L, 0.01);
          s_lim[g][1] = percentile_sorted(s_sort, L, 0.99);
        }
        __syncthreads();
      }
      for (int g = 0; g < 2; ++g)
        for (int e = threadIdx.x; e < L; e += 256) s_v[g][e] = fminf(fmaxf(s_v[g][e], s_lim[g][0]), s_lim[g][1]);
      __syncthreads();
    }
    // ---- Pearson correlation (np.corrcoef) ----
    double sa = 0.0, sb = 0.0;
    for (int e = threadIdx.x; e < L; e += 256) { sa += (double)s_v[0][e]; sb += (double)s_v[1][e]; }
    const double ma = block_sum_d(sa, s_red) / L, mb = block_sum_d(sb, s_red) / L;
    double saa = 0.0, sbb = 0.0, sab = 0.0;
    for (int e = threadIdx.x; e < L; e += 256) {
      const double da = (double)s_v[0][e] - ma, db = (double)s_v[1][e] - mb;
      saa += da * da; sbb += db * db; sab += da * db;
    }
    saa = block_sum_d(saa, s_red); sbb = block_sum_d(sbb, s_red); sab = block_sum_d(sab, s_red);
    if (threadIdx.x == 0) {
      const double den = sqrt(saa) * sqrt(sbb);
      double c = sab / den;  // NaN when a clipped row is constant, like np.corrcoef
      const bool keep = valid && !(c != c) && !isinf(c);
      if (keep) c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
      val[(int64_t)blockIdx.x * 4 + metric] = keep ? (float)c : 0.f;
      ok[(int64_t)blockIdx.x * 4 + metric] = keep ? 1.f : 0.f;
    }
    __syncthreads();
  }
}

// out[m] = sum_i val[i][m] / sum_i ok[i][m]  (0 if nothing was kept); one block, fixed order
__global__ void __launch_bounds__(256) dynamics_finalize_kernel(const float* __restrict__ val, const float* __restrict__ ok, int n,
                                                                float* __restrict__ out) {
  __shared__ double s_red[256];
  for (int m = 0; m < 4; ++m) {
    double sv = 0.0, so = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { sv += (double)val[(int64_t)i * 4 + m]; so += (double)ok[(int64_t)i * 4 + m]; }
    sv = block_sum_d(sv, s_red);
    so = block_sum_d(so, s_red);
    if (threadIdx.x == 0) out[m] = so > 0.0 ? (float)(sv / so) : 0.f;
  }
}

// out[0] = mean(x[0..n))  in double, one block, fixed order
__global__ void __launch_bounds__(256) mean_d_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  __shared__ double s_red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)x[i];
  s = block_sum_d(s, s_red);
  if (threadIdx.x == 0) out[0] = (float)(s / n);
}

}  // namespace ev

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int wgg_eval_cdist(wgg_ctx* ctx, const float* a, int64_t na, const float* b, int64_t nb, int32_t d, float* out,
                              void* stream) {
  if (!ctx || !a || !b || !out || na < 0 || nb < 0 || d <= 0) return wgg_fail(ctx, WGG_EINVAL, "wgg_eval_cdist: bad argument%s");
  if (na == 0 || nb == 0) return WGG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)cdiv64(nb, 32), (unsigned)cdiv64(na, 32));
  ev::cdist_kernel<<<grid, 256, 0, st>>>(a, (int)na, b, (int)nb, d, out);
  WGG_CHECK_LAUNCH(ctx, "cdist_kernel");
  return WGG_OK;
}

extern "C" int wgg_eval_row_kth(wgg_ctx* ctx, const float* m, int64_t rows, int64_t cols, int32_t k, float* out, void* stream) {
  if (!ctx || !m || !out || rows < 0 || cols <= 0 || k < 0 || k >= ev::KTH_MAX || k >= cols)
    return wgg_fail(ctx, WGG_EINVAL, "wgg_eval_row_kth: bad argument (k must be < 8 and < cols)%s");
  if (rows == 0) return WGG_OK;
  ev::row_kth_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(m, (int)cols, k, out);
  WGG_CHECK_LAUNCH(ctx, "row_kth_kernel");
  return WGG_OK;
}

extern "C" int wgg_eval_precision_recall(wgg_ctx* ctx, const float* rf, int64_t n_real, int64_t n_fake, const float* real_radii,
                                         const float* fake_radii, float* out2, float* ws2, void* stream) {
  if (!ctx || !rf || !real_radii || !fake_radii || !out2 || !ws2 || n_real <= 0 || n_fake <= 0)
    return wgg_fail(ctx, WGG_EINVAL, "wgg_eval_precision_recall: bad argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(ws2, 0, 2 * sizeof(float), st) != cudaSuccess) return wgg_fail(ctx, WGG_ECUDA, "wgg_eval_precision_recall: memset%s");
  ev::prec_kernel<<<(unsigned)cdiv64(n_fake, 256), 256, 0, st>>>(rf, (int)n_real, (int)n_fake, real_radii, ws2);
  WGG_CHECK_LAUNCH(ctx, "prec_kernel");
  ev::rec_kernel<<<(unsigned)n_real, 128, 0, st>>>(rf, (int)n_fake, fake_radii, ws2);
  WGG_CHECK_LAUNCH(ctx, "rec_kernel");
  ev::prec_rec_finalize_kernel<<<1, 1, 0, st>>>(ws2, (int)n_real, (int)n_fake, out2);
  WGG_CHECK_LAUNCH(ctx, "prec_rec_finalize_kernel");
  return WGG_OK;
}

extern "C" int wgg_eval_jerk(wgg_ctx* ctx, const float* g, int64_t n, int32_t T, int32_t C, const float* S, float* out, float* ws,
                             void* stream) {
  if (!ctx || !g || !S || !out || !ws || n <= 0 || T <= 0 || C < 2) return wgg_fail(ctx, WGG_EINVAL, "wgg_eval_jerk: bad argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  ev::jerk_kernel<<<(unsigned)n, 128, (size_t)(2 * T) * sizeof(float), st>>>(g, T, C, S, ws);
  WGG_CHECK_LAUNCH(ctx, "jerk_kernel");
  ev::mean_d_kernel<<<1, 256, 0, st>>>(ws, (int)n, out);
  WGG_CHECK_LAUNCH(ctx, "mean_d_kernel");
  return WGG_OK;
}

extern "C" int wgg_eval_dynamics(wgg_ctx* ctx, const float* real, const float* fake, int64_t n, int32_t T, int32_t C, float* out4,
                                 float* ws, void* stream) {
  if (!ctx || !real || !fake || !out4 || !ws || n <= 0 || T < 3 || C < 3)
    return wgg_fail(ctx, WGG_EINVAL, "wgg_eval_dynamics: bad argument%s");
  if (2 * (T - 1) > ev::DYN_MAX) return wgg_fail(ctx, WGG_EUNSUPPORTED, "wgg_eval_dynamics: seq_length > 257%s");
  cudaStream_t st = (cudaStream_t)stream;
  ev::dynamics_kernel<<<(unsigned)n, 256, 0, st>>>(real, fake, T, C, ws, ws + 4 * n);
  WGG_CHECK_LAUNCH(ctx, "dynamics_kernel");
  ev::dynamics_finalize_kernel<<<1, 256, 0, st>>>(ws, ws + 4 * n, (int)n, out4);
  WGG_CHECK_LAUNCH(ctx, "dynamics_finalize_kernel");
  return WGG_OK;
}

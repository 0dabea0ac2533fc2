// Word prototypes and minimum-jerk trajectories on the GPU: the geometry of
// src/shared/keyboard.py:389-514 (generate_minimum_jerk_trajectory) and :710-765 (QWERTYKeyboard.get_word_prototype),
// one thread block per word, float64 arithmetic like numpy, float32 (x, y, t) rows out.  Random draws stay with the
// caller (np.random.normal in the reference): the key-offset and midpoint noise are inputs.
#include <math.h>

#include "common.cuh"

namespace kb {

constexpr int MAXP = 64;     // via-points per word after midpoint expansion (2 k - 1 for k keys: k <= 32)
constexpr int NFINE = 1000;  // fine tau resolution (keyboard.py:488)

// numpy.linspace(0, stop, n)[i]
__device__ __forceinline__ double linspace0(double stop, int n, int i) {
  if (n == 1) return 0.0;
  return i == n - 1 ? stop : (double)i * (stop / (double)(n - 1));
}

// keys (n, maxk, 2) float64 key centres, nkeys (n) -> out (n, T, 3): straight segments sampled at uniform arc length,
// t = linspace(0, 1)   (keyboard.py:710-765; single-point / empty cases :688-694, :726-729, :741-743)
__global__ void __launch_bounds__(128) prototype_kernel(const double* __restrict__ keys, const int* __restrict__ nkeys, int maxk,
                                                        int T, float* __restrict__ out) {
  __shared__ double sx[MAXP], sy[MAXP], cum[MAXP], seg[MAXP];
  const int w = blockIdx.x;
  const int k = min(nkeys[w], maxk);
  const double* kp = keys + (int64_t)w * maxk * 2;
  for (int i = threadIdx.x; i < k; i += 128) { sx[i] = kp[2 * i]; sy[i] = kp[2 * i + 1]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    cum[0] = 0.0;
    for (int i = 0; i + 1 < k; ++i) {
      const double dx = sx[i + 1] - sx[i], dy = sy[i + 1] - sy[i];
      seg[i] = sqrt(dx * dx + dy * dy);   // np.linalg.norm
      cum[i + 1] = cum[i] + seg[i];       // np.cumsum, sequential like numpy
    }
  }
  __syncthreads();
  float* o = out + (int64_t)w * T * 3;
  const double total = k >= 2 ? cum[k - 1] : 0.0;
  for (int i = threadIdx.x; i < T; i += 128) {
    float x = 0.f, y = 0.f;
    if (k == 1 || (k >= 2 && total < 1e-6)) {
      x = (float)sx[0]; y = (float)sy[0];
    } else if (k >= 2) {
      const double target = linspace0(total, T, i);
      // np.searchsorted(cum, target, side='right') - 1, clamped to [0, k - 2]
      int s = 0;
      while (s < k && cum[s] <= target) ++s;
      s = max(0, min(s - 1, k - 2));
      double t = seg[s] > 1e-6 ? (target - cum[s]) / seg[s] : 0.0;
      t = fmax(0.0, fmin(1.0, t));
      x = (float)(sx[s] + t * (sx[s + 1] - sx[s]));
      y = (float)(sy[s] + t * (sy[s + 1] - sy[s]));
    }
    o[i * 3] = x;
    o[i * 3 + 1] = y;
    o[i * 3 + 2] = k == 0 ? 0.f : (float)linspace0(1.0, T, i);   // a word without any known key: all zeros (:727-729)
  }
}

// numpy.interp(x, xp, fp) for non-decreasing xp of length n
__device__ double interp1(double x, const double* xp, const double* fp, int n) {
  if (x <= xp[0]) return fp[0];
  if (x >= xp[n - 1]) return fp[n - 1];
  int lo = 0, hi = n - 1;  // invariant xp[lo] <= x < xp[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid; else hi = mid;
  }
  const double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
  return slope * (x - xp[lo]) + fp[lo];
}

// via (n, maxk, 2) key centres, nkeys (n), key_noise (n, maxk, 2) [rows 1 .. k-2 used], mid_noise (n, maxk) [k-1 used]
// (both already scaled: N(0, offset_std) and N(0, offset_std / 2); pass zeros for offset_std = 0) -> out (n, T, 3)
__global__ void __launch_bounds__(256) minjerk_kernel(const double* __restrict__ via, const int* __restrict__ nkeys, int maxk,
                                                      const double* __restrict__ key_noise, const double* __restrict__ mid_noise,
                                                      int include_midpoints, int apply_noise, int T, float* __restrict__ out) {
  __shared__ double px[MAXP], py[MAXP], vx[MAXP], vy[MAXP];
  __shared__ double fx[NFINE], fy[NFINE], fs[NFINE];
  __shared__ int s_n;
  const int w = blockIdx.x;
  const int k = min(nkeys[w], maxk);
  const double* vp = via + (int64_t)w * maxk * 2;
  float* o = out + (int64_t)w * T * 3;
  if (k < 2) {  // keyboard.py:846-851: one key -> that point with uniform time; no key -> all zeros
    for (int i = threadIdx.x; i < T; i += 256) {
      o[i * 3] = k == 1 ? (float)vp[0] : 0.f;
      o[i * 3 + 1] = k == 1 ? (float)vp[1] : 0.f;
      o[i * 3 + 2] = k == 1 ? (float)linspace0(1.0, T, i) : 0.f;
    }
    return;
  }
  if (threadIdx.x == 0) {
    // offset noise on the interior key centres (:426-429), then midpoints with perpendicular noise (:432-445)
    double qx[MAXP / 2 + 1], qy[MAXP / 2 + 1];
    for (int i = 0; i < k; ++i) {
      qx[i] = vp[2 * i]; qy[i] = vp[2 * i + 1];
      if (apply_noise && k > 2 && i >= 1 && i <= k - 2) {
        qx[i] += key_noise[((int64_t)w * maxk + (i - 1)) * 2];
        qy[i] += key_noise[((int64_t)w * maxk + (i - 1)) * 2 + 1];
      }
    }
    int n = 0;
    if (include_midpoints && k > 2) {
      px[n] = qx[0]; py[n] = qy[0]; ++n;
      for (int i = 0; i + 1 < k; ++i) {
        double mx = (qx[i] + qx[i + 1]) / 2, my = (qy[i] + qy[i + 1]) / 2;
        if (apply_noise) {
          const double dx = qx[i + 1] - qx[i], dy = qy[i + 1] - qy[i];
          double ex = -dy, ey = dx;
          const double nrm = sqrt(ex * ex + ey * ey) + 1e-8;
          ex /= nrm; ey /= nrm;
          const double z = mid_noise[(int64_t)w * maxk + i];
          mx += ex * z; my += ey * z;
        }
        px[n] = mx; py[n] = my; ++n;
        px[n] = qx[i + 1]; py[n] = qy[i + 1]; ++n;
      }
    } else {
      for (int i = 0; i < k; ++i) { px[i] = qx[i]; py[i] = qy[i]; }
      n = k;
    }
    s_n = n;
    // Catmull-Rom style tangents scaled by the harmonic mean of the neighbouring segment lengths (:462-476)
    for (int i = 0; i < n; ++i) { vx[i] = 0.0; vy[i] = 0.0; }
    for (int i = 1; i + 1 < n; ++i) {
      const double bx = px[i] - px[i - 1], by = py[i] - py[i - 1], ax = px[i + 1] - px[i], ay = py[i + 1] - py[i];
      const double lb = sqrt(bx * bx + by * by), la = sqrt(ax * ax + ay * ay);
      if (lb > 1e-6 && la > 1e-6) {
        const double tx = (bx / lb + ax / la) / 2, ty = (by / lb + ay / la) / 2;
        const double scale = 2 * lb * la / (lb + la);
        vx[i] = tx * scale; vy[i] = ty * scale;
      }
    }
  }
  __syncthreads();
  const int n = s_n;
  if (n == 2) {  // one segment, quintic profile, time = tau (:449-455)
    for (int i = threadIdx.x; i < T; i += 256) {
      const double tau = linspace0(1.0, T, i);
      const double t3 = tau * tau * tau;
      const double s = 10 * t3 - 15 * t3 * tau + 6 * t3 * tau * tau;
      o[i * 3] = (float)(px[0] + s * (px[1] - px[0]));
      o[i * 3 + 1] = (float)(py[0] + s * (py[1] - py[0]));
      o[i * 3 + 2] = (float)tau;
    }
    return;
  }
  // fine trajectory: quintic Hermite segments with zero accelerations (:341-386, :295-338)
  for (int i = threadIdx.x; i < NFINE; i += 256) {
    const double tau = linspace0(1.0, NFINE, i);
    const double st = tau * (n - 1);
    int sidx = (int)st;
    if (sidx > n - 2) sidx = n - 2;
    const double t = st - sidx;
    const double t2 = t * t, t3 = t2 * t, t4 = t3 * t, t5 = t4 * t;
    const double h00 = 1 - 10 * t3 + 15 * t4 - 6 * t5, h01 = 10 * t3 - 15 * t4 + 6 * t5;
    const double h10 = t - 6 * t3 + 8 * t4 - 3 * t5, h11 = -4 * t3 + 7 * t4 - 3 * t5;
    // np.outer sums left to right: h00 p0 + h01 p1 + h10 v0 + h11 v1 (+ 0 + 0)
    fx[i] = ((h00 * px[sidx] + h01 * px[sidx + 1]) + h10 * vx[sidx]) + h11 * vx[sidx + 1];
    fy[i] = ((h00 * py[sidx] + h01 * py[sidx + 1]) + h10 * vy[sidx]) + h11 * vy[sidx + 1];
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // s(tau): sequential cumulative sum like np.cumsum (:492-494)
    fs[0] = 0.0;
    for (int i = 1; i < NFINE; ++i) {
      const double dx = fx[i] - fx[i - 1], dy = fy[i] - fy[i - 1];
      fs[i] = fs[i - 1] + sqrt(dx * dx + dy * dy);
    }
  }
  __syncthreads();
  const double total = fs[NFINE - 1];
  for (int i = threadIdx.x; i < T; i += 256) {
    if (total < 1e-6) {  // degenerate (:496-500)
      o[i * 3] = (float)px[0]; o[i * 3 + 1] = (float)py[0]; o[i * 3 + 2] = (float)linspace0(1.0, T, i);
      continue;
    }
    const double st = linspace0(total, T, i);
    // tau_fine is linspace(0, 1, NFINE): interpolate the index, then map (:506-511)
    double lo_tau;
    {
      // np.interp(s_target, s_of_tau, tau_fine)
      if (st <= fs[0]) lo_tau = 0.0;
      else if (st >= fs[NFINE - 1]) lo_tau = 1.0;
      else {
        int lo = 0, hi = NFINE - 1;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (fs[mid] <= st) lo = mid; else hi = mid; }
        const double t0 = linspace0(1.0, NFINE, lo), t1 = linspace0(1.0, NFINE, lo + 1);
        lo_tau = (t1 - t0) / (fs[lo + 1] - fs[lo]) * (st - fs[lo]) + t0;
      }
    }
    o[i * 3] = (float)interp1(st, fs, fx, NFINE);
    o[i * 3 + 1] = (float)interp1(st, fs, fy, NFINE);
    o[i * 3 + 2] = (float)lo_tau;
  }
}

}  // namespace kb

extern "C" int wgg_word_prototypes(wgg_ctx* ctx, const double* keys, const int32_t* nkeys, int64_t n, int32_t maxk, int32_t T,
                                   float* out, void* stream) {
  if (!ctx || !keys || !nkeys || !out || n < 0 || maxk < 1 || maxk > kb::MAXP || T < 1)
    return wgg_fail(ctx, WGG_EINVAL, "wgg_word_prototypes: bad argument (at most 64 keys per word)%s");
  if (n == 0) return WGG_OK;
  kb::prototype_kernel<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(keys, nkeys, maxk, T, out);
  WGG_CHECK_LAUNCH(ctx, "prototype_kernel");
  return WGG_OK;
}

extern "C" int wgg_minimum_jerk(wgg_ctx* ctx, const double* keys, const int32_t* nkeys, int64_t n, int32_t maxk,
                                const double* key_noise, const double* mid_noise, int include_midpoints, int32_t T, float* out,
                                void* stream) {
  if (!ctx || !keys || !nkeys || !out || n < 0 || maxk < 1 || 2 * maxk - 1 > kb::MAXP || T < 1)
    return wgg_fail(ctx, WGG_EINVAL, "wgg_minimum_jerk: bad argument (at most 32 keys per word)%s");
  if ((key_noise == nullptr) != (mid_noise == nullptr)) return wgg_fail(ctx, WGG_EINVAL, "wgg_minimum_jerk: give both noise arrays or none%s");
  if (n == 0) return WGG_OK;
  kb::minjerk_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(keys, nkeys, maxk, key_noise, mid_noise, include_midpoints,
                                                                     key_noise != nullptr, T, out);
  WGG_CHECK_LAUNCH(ctx, "minjerk_kernel");
  return WGG_OK;
}

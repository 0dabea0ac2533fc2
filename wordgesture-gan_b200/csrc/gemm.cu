// Generic strided contraction engine used by every dense layer / conv / weight-gradient of the
// WordGesture-GAN step (src/gan/models.py nn.Linear / nn.Conv1d call sites and their autograd).
//
//   C(m,n) = act( sum_k A(m,k) * B(k,n) + bias[n] + bias2[n] )
//
// One kernel serves forward (NT), backward-data (NN) and backward-weight (TN, deterministic
// split-K) forms through operand strides; conv1d is an implicit GEMM through a sliding-window
// view of the channel-last activation tensor (no im2col materialisation).
// fp32 FMA path: 64x64x16 tiles, 256 threads, 4x4 register tile, register-prefetched global loads.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <int CONV>
__global__ void __launch_bounds__(256) gemm_kernel(GemmP p) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int batch = blockIdx.z / p.splitk, split = blockIdx.z % p.splitk;
  const float* __restrict__ A = p.A + (int64_t)batch * p.bsA;
  const float* __restrict__ B = p.B + (int64_t)batch * p.bsB;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const int64_t ktiles = (p.K + BK - 1) / BK;
  const int64_t per = (ktiles + p.splitk - 1) / p.splitk;
  const int64_t kt_begin = (int64_t)split * per;
  const int64_t kt_end = min(ktiles, kt_begin + per);

  // thread -> tile element maps chosen so that global reads run along the unit-stride dimension
  const bool a_kfast = (p.sak == 1);
  const bool b_nfast = (p.sbn == 1) || (p.sbk != 1);
  int a_mm[4], a_kk[4], b_kk[4], b_nn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * 256;
    if (a_kfast) { a_kk[i] = idx % BK; a_mm[i] = idx / BK; } else { a_mm[i] = idx % BM; a_kk[i] = idx / BM; }
    if (b_nfast) { b_nn[i] = idx % BN; b_kk[i] = idx / BN; } else { b_kk[i] = idx % BK; b_nn[i] = idx / BK; }
  }

  auto load_a = [&](int64_t k0, float (&r)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + a_mm[i], k = k0 + a_kk[i];
      float v = 0.f;
      if (m < p.M && k < p.K) {
        bool ok = true;
        if (CONV == 1) {
          const int tt = (int)(m % p.conv_T) + (int)(k / p.conv_Cin) - p.conv_pad;
          ok = (unsigned)tt < (unsigned)p.conv_T;
        }
        if (ok) v = __ldg(A + m * p.sam + k * p.sak);
      }
      r[i] = v;
    }
  };
  auto load_b = [&](int64_t k0, float (&r)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t k = k0 + b_kk[i], n = n0 + b_nn[i];
      float v = 0.f;
      if (k < p.K && n < p.N) {
        bool ok = true;
        if (CONV == 2) {
          const int tt = (int)(k % p.conv_T) + (int)(n / p.conv_Cin) - p.conv_pad;
          ok = (unsigned)tt < (unsigned)p.conv_T;
        }
        if (ok) v = __ldg(B + k * p.sbk + n * p.sbn);
      }
      r[i] = v;
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int tx = tid % 16, ty = tid / 16;
  float ra[4], rb[4];
  if (kt_begin < kt_end) {
    load_a(kt_begin * BK, ra);
    load_b(kt_begin * BK, rb);
  }
  for (int64_t kt = kt_begin; kt < kt_end; ++kt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[a_kk[i]][a_mm[i]] = ra[i];
      Bs[b_kk[i]][b_nn[i]] = rb[i];
    }
    __syncthreads();
    if (kt + 1 < kt_end) {
      load_a((kt + 1) * BK, ra);
      load_b((kt + 1) * BK, rb);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (p.splitk > 1) {
    float* P = p.partial + (int64_t)blockIdx.z * p.M * p.N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + ty * 4 + i;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + tx * 4 + j;
        if (n < p.N) P[m * p.N + n] = acc[i][j];
      }
    }
    return;
  }
  float* C = p.C + (int64_t)batch * p.bsC;
  const float* bias = p.bias ? p.bias + (int64_t)batch * p.bsBias : nullptr;
  const float* bias2 = p.bias2 ? p.bias2 + (int64_t)batch * p.bsBias : nullptr;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t n = n0 + tx * 4 + j;
    if (n >= p.N) continue;
    float bsum = 0.f;
    if (bias) bsum += __ldg(bias + n);
    if (bias2) bsum += __ldg(bias2 + n);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + ty * 4 + i;
      if (m >= p.M) continue;
      float v = acc[i][j] + bsum;
      if (p.act == ACT_LEAKY) v = leaky_f(v);
      else if (p.act == ACT_TANH) v = tanhf(v);
      float* dst = C + m * p.scm + n * p.scn;
      if (p.accumulate) v += *dst;
      *dst = v;
    }
  }
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ P, int S, int64_t n,
                                                              int64_t bsP, float* __restrict__ out,
                                                              float* __restrict__ out2, int64_t bsOut,
                                                              int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = blockIdx.y;
  const float* p = P + (int64_t)b * bsP + i;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += p[(int64_t)k * n];
  float* o = out + (int64_t)b * bsOut + i;
  *o = accumulate ? *o + s : s;
  if (out2) {
    float* o2 = out2 + (int64_t)b * bsOut + i;
    *o2 = accumulate ? *o2 + s : s;
  }
}

// stage 1 of the column sum: block (32 cols, 8 row-lanes); grid (ceil(N/32), S, nbatch)
__global__ void __launch_bounds__(256) colsum_stage1(const float* __restrict__ X, int64_t M, int64_t N, int64_t ldx,
                                                     int64_t bsX, int S, float* __restrict__ P) {
  __shared__ float sh[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * 32 + tx;
  const int s = blockIdx.y, b = blockIdx.z;
  const int64_t rows_per = (M + S - 1) / S;
  const int64_t r0 = (int64_t)s * rows_per, r1 = min(M, r0 + rows_per);
  const float* x = X + (int64_t)b * bsX;
  float acc = 0.f;
  if (n < N)
    for (int64_t r = r0 + ty; r < r1; r += 8) acc += __ldg(x + r * ldx + n);
  sh[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][tx];
    P[((int64_t)b * S + s) * N + n] = t;
  }
}

constexpr int kMaxSplit = 128;

}  // namespace

int gemm_choose_splitk(wgg_ctx* ctx, int64_t M, int64_t N, int64_t K, int nbatch) {
  const int64_t tiles = cdiv64(M, BM) * cdiv64(N, BN) * nbatch;
  const int64_t ktiles = cdiv64(K, BK);
  int64_t want = cdiv64(2 * (int64_t)ctx->sm_count, tiles);
  int64_t maxs = ktiles / 8;  // at least 8 k-tiles per split
  if (want > maxs) want = maxs;
  if (want > kMaxSplit) want = kMaxSplit;
  if (want < 1) want = 1;
  return (int)want;
}

int64_t gemm_splitk_ws_floats(int64_t M, int64_t N, int nbatch) { return (int64_t)kMaxSplit * M * N * nbatch; }

int gemm_launch(wgg_ctx* ctx, const GemmP& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return WGG_OK;
  if (p.M >= (1ll << 31) || p.K >= (1ll << 31)) return wgg_fail(ctx, WGG_EINVAL, "gemm: dimension too large%s");
  if (p.splitk > 1 && (!p.partial || p.act != ACT_NONE || p.bias || p.bias2))
    return wgg_fail(ctx, WGG_EINVAL, "gemm: split-K needs a partial buffer and a plain epilogue%s");
  dim3 grid((unsigned)cdiv64(p.M, BM), (unsigned)cdiv64(p.N, BN), (unsigned)(p.nbatch * p.splitk));
  if (grid.y > 65535 || grid.z > 65535) return wgg_fail(ctx, WGG_EINVAL, "gemm: grid too large%s");
  ProfScope prof(ctx, "gemm_kernel", st, 2.0 * (double)p.M * (double)p.N * (double)p.K * p.nbatch,
                 4.0 * ((double)p.M * p.K + (double)p.K * p.N + (double)p.M * p.N) * p.nbatch);
  if (p.conv_mode == 1) gemm_kernel<1><<<grid, 256, 0, st>>>(p);
  else if (p.conv_mode == 2) gemm_kernel<2><<<grid, 256, 0, st>>>(p);
  else gemm_kernel<0><<<grid, 256, 0, st>>>(p);
  WGG_CHECK_LAUNCH(ctx, "gemm_kernel");
  if (p.splitk > 1) {
    if (p.scn != 1 || p.scm != p.N)
      return wgg_fail(ctx, WGG_EINVAL, "gemm: split-K output must be dense row-major%s");
    WGG_TRY(reduce_partials_launch(ctx, p.partial, p.splitk, p.M * p.N, p.nbatch, (int64_t)p.splitk * p.M * p.N, p.C,
                                   nullptr, p.bsC, p.accumulate, st));
  }
  return WGG_OK;
}

int reduce_partials_launch(wgg_ctx* ctx, const float* P, int S, int64_t n, int nbatch, int64_t bsP, float* out,
                           float* out2, int64_t bsOut, int accumulate, cudaStream_t st) {
  if (n <= 0) return WGG_OK;
  dim3 grid((unsigned)cdiv64(n, 256), (unsigned)nbatch);
  reduce_partials_kernel<<<grid, 256, 0, st>>>(P, S, n, bsP, out, out2, bsOut, accumulate);
  WGG_CHECK_LAUNCH(ctx, "reduce_partials_kernel");
  return WGG_OK;
}

int64_t colsum_ws_floats(int64_t N, int nbatch) { return (int64_t)kMaxSplit * N * nbatch; }

int colsum_launch(wgg_ctx* ctx, const float* X, int64_t M, int64_t N, int64_t ldx, int nbatch, int64_t bsX,
                  float* out, float* out2, int64_t bsOut, int accumulate, float* ws, cudaStream_t st) {
  if (N <= 0) return WGG_OK;
  int64_t S = cdiv64(M, 256);
  const int64_t colblocks = cdiv64(N, 32) * nbatch;
  const int64_t want = cdiv64(4 * (int64_t)ctx->sm_count, colblocks);
  if (S > want) S = want;
  if (S > kMaxSplit) S = kMaxSplit;
  if (S < 1) S = 1;
  dim3 grid((unsigned)cdiv64(N, 32), (unsigned)S, (unsigned)nbatch);
  colsum_stage1<<<grid, 256, 0, st>>>(X, M, N, ldx, bsX, (int)S, ws);
  WGG_CHECK_LAUNCH(ctx, "colsum_stage1");
  return reduce_partials_launch(ctx, ws, (int)S, N, nbatch, S * N, out, out2, bsOut, accumulate, st);
}

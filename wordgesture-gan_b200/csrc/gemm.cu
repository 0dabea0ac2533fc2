// Generic strided contraction engine used by every dense layer / conv / weight-gradient of the
// WordGesture-GAN step (src/gan/models.py nn.Linear / nn.Conv1d call sites and their autograd).
//
//   C(m,n) = act( sum_k A(m,k) * B(k,n) + bias[n] + bias2[n] )
//
// One kernel serves forward (NT), backward-data (NN) and backward-weight (TN, deterministic
// split-K) forms through operand strides; conv1d is an implicit GEMM through a sliding-window
// view of the channel-last activation tensor (no im2col materialisation).
// fp32 FMA path: 64x64x16 tiles, 256 threads, 4x4 register tile, register-prefetched global loads.
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;     // fp32 FMA kernel tile
constexpr int TBM = 128, TBN = 64, TBK = 16, TPAD = 8;  // tensor-core kernel tile

__device__ __forceinline__ int fast_div(int x, int d, float inv) {
  int q = (int)(((float)x + 0.5f) * inv);
  const int r = x - q * d;  // one correction step: exact for 0 <= x < 2^23
  q += (r >= d) - (r < 0);
  return q;
}

// ---------------------------------------------------------------------------------------------
// Operand tile loaders.  Every division / modulo is hoisted out of the K loop: per thread the (m | n)
// coordinate of each element is fixed and only k advances by the tile depth, so global offsets are advanced
// incrementally and the conv-window test costs one add + one compare per element.
// ---------------------------------------------------------------------------------------------
// Element e of a thread sits at tile coordinates (c0 + e*dc) in the fixed (m | n) direction and (k0 + e*dk) in
// k, an affine pattern, so only the bases live in registers.
template <int NE, int TILE_MN, int TILE_K, int CONV>
struct ALoader {
  int64_t off0, doff, kstep;
  int mm0, kk0, dmm, dkk;
  int mt[CONV == 1 ? NE : 1];
  float inv_cin;
  __device__ __forceinline__ int mm(int e) const { return mm0 + e * dmm; }
  __device__ __forceinline__ int kk(int e) const { return kk0 + e * dkk; }
  __device__ __forceinline__ void init(const GemmP& p, int64_t m0, int64_t k_begin, int tid) {
    if (p.sak == 1) { kk0 = tid % TILE_K; mm0 = tid / TILE_K; dkk = 0; dmm = 256 / TILE_K; }
    else { mm0 = tid % TILE_MN; kk0 = tid / TILE_MN; dmm = 0; dkk = 256 / TILE_MN; }
    kstep = (int64_t)TILE_K * p.sak;
    off0 = (m0 + mm0) * p.sam + (k_begin + kk0) * p.sak;
    doff = (int64_t)dmm * p.sam + (int64_t)dkk * p.sak;
    inv_cin = CONV == 1 ? 1.f / (float)p.conv_Cin : 0.f;
    if (CONV == 1) {
#pragma unroll
      for (int e = 0; e < NE; ++e) mt[e] = (int)((m0 + mm(e)) % p.conv_T) - p.conv_pad;
    }
  }
  __device__ __forceinline__ void load(const GemmP& p, const float* __restrict__ A, int64_t m0, int64_t k0,
                                       float (&r)[NE]) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int64_t k = k0 + kk(e);
      bool v = (m0 + mm(e) < p.M) && k < p.K;
      if (CONV == 1 && v) v = (unsigned)(mt[e] + fast_div((int)k, p.conv_Cin, inv_cin)) < (unsigned)p.conv_T;
      r[e] = v ? __ldg(A + off0 + e * doff) : 0.f;
    }
    off0 += kstep;
  }
};

template <int NE, int TILE_MN, int TILE_K, int CONV>
struct BLoader {
  int64_t off0, doff, kstep;
  int nn0, kk0, dnn, dkk;
  int nc[CONV == 2 ? NE : 1], kmod[CONV == 2 ? NE : 1];
  __device__ __forceinline__ int nn(int e) const { return nn0 + e * dnn; }
  __device__ __forceinline__ int kk(int e) const { return kk0 + e * dkk; }
  __device__ __forceinline__ void init(const GemmP& p, int64_t n0, int64_t k_begin, int tid) {
    if ((p.sbn == 1) || (p.sbk != 1)) { nn0 = tid % TILE_MN; kk0 = tid / TILE_MN; dnn = 0; dkk = 256 / TILE_MN; }
    else { kk0 = tid % TILE_K; nn0 = tid / TILE_K; dkk = 0; dnn = 256 / TILE_K; }
    kstep = (int64_t)TILE_K * p.sbk;
    off0 = (k_begin + kk0) * p.sbk + (n0 + nn0) * p.sbn;
    doff = (int64_t)dkk * p.sbk + (int64_t)dnn * p.sbn;
    if (CONV == 2) {
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        nc[e] = (int)((n0 + nn(e)) / p.conv_Cin) - p.conv_pad;
        kmod[e] = (int)((k_begin + kk(e)) % p.conv_T);
      }
    }
  }
  __device__ __forceinline__ void load(const GemmP& p, const float* __restrict__ B, int64_t n0, int64_t k0,
                                       float (&r)[NE]) {
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int64_t k = k0 + kk(e);
      bool v = (n0 + nn(e) < p.N) && k < p.K;
      if (CONV == 2) {
        v = v && (unsigned)(kmod[e] + nc[e]) < (unsigned)p.conv_T;
        kmod[e] += TILE_K;
        while (kmod[e] >= p.conv_T) kmod[e] -= p.conv_T;
      }
      r[e] = v ? __ldg(B + off0 + e * doff) : 0.f;
    }
    off0 += kstep;
  }
};

__device__ __forceinline__ void store_out(const GemmP& p, float* C, const float* bias, const float* bias2, int64_t m,
                                          int64_t n, float v) {
  if (bias) v += __ldg(bias + n);
  if (bias2) v += __ldg(bias2 + n);
  if (p.act == ACT_LEAKY) v = leaky_f(v);
  else if (p.act == ACT_TANH) v = tanhf(v);
  float* dst = C + m * p.scm + n * p.scn;
  if (p.accumulate) v += *dst;
  *dst = v;
}

// ---------------------------------------------------------------------------------------------
// fp32 FMA kernel: 64x64x16 tile, 256 threads, 4x4 register tile
// ---------------------------------------------------------------------------------------------
template <int CONV, int BM_>
__global__ void __launch_bounds__(256, BM_ == 64 ? 2 : 4) gemm_kernel(GemmP p) {
  constexpr int RM = BM_ / 16;  // rows per thread (4, or 2 for the small-problem tile that doubles the CTA count)
  __shared__ __align__(16) float As[BK][BM_ + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int batch = blockIdx.z / p.splitk, split = blockIdx.z % p.splitk;
  const float* __restrict__ A = p.A + (int64_t)batch * p.bsA;
  const float* __restrict__ B = p.B + (int64_t)batch * p.bsB;
  const int64_t m0 = (int64_t)blockIdx.x * BM_;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const int64_t ktiles = (p.K + BK - 1) / BK;
  const int64_t per = (ktiles + p.splitk - 1) / p.splitk;
  const int64_t kt_begin = (int64_t)split * per;
  const int64_t kt_end = min(ktiles, kt_begin + per);

  ALoader<RM, BM_, BK, CONV> la;
  BLoader<4, BN, BK, CONV> lb;
  la.init(p, m0, kt_begin * BK, tid);
  lb.init(p, n0, kt_begin * BK, tid);

  float acc[RM][4], rs[RM];
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    rs[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }

  const int tx = tid % 16, ty = tid / 16;
  float ra[RM], rb[4];
  if (kt_begin < kt_end) {
    la.load(p, A, m0, kt_begin * BK, ra);
    lb.load(p, B, n0, kt_begin * BK, rb);
  }
  for (int64_t kt = kt_begin; kt < kt_end; ++kt) {
#pragma unroll
    for (int i = 0; i < RM; ++i) As[la.kk(i)][la.mm(i)] = ra[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[lb.kk(i)][lb.nn(i)] = rb[i];
    __syncthreads();
    if (kt + 1 < kt_end) {
      la.load(p, A, m0, (kt + 1) * BK, ra);
      lb.load(p, B, n0, (kt + 1) * BK, rb);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[RM];
      if (RM == 4) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        av[0] = a.x; av[1] = a.y; av[RM - 2] = a.z; av[RM - 1] = a.w;
      } else {
        const float2 a = *reinterpret_cast<const float2*>(&As[kk][ty * 2]);
        av[0] = a.x; av[1] = a.y;
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        rs[i] += av[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }

  const int64_t pstride = p.M * p.N + (p.rowsum ? p.M : 0);
  float* P = p.splitk > 1 ? p.partial + (int64_t)blockIdx.z * pstride : nullptr;
  if (p.rowsum && blockIdx.y == 0 && tx == 0) {
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int64_t m = m0 + ty * RM + i;
      if (m < p.M) {
        if (P) P[p.M * p.N + m] = rs[i];
        else p.rowsum[m] += rs[i];
      }
    }
  }
  float* C = p.C + (int64_t)batch * p.bsC;
  const float* bias = p.bias ? p.bias + (int64_t)batch * p.bsBias : nullptr;
  const float* bias2 = p.bias2 ? p.bias2 + (int64_t)batch * p.bsBias : nullptr;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int64_t m = m0 + ty * RM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      if (P) P[m * p.N + n] = acc[i][j];
      else store_out(p, C, bias, bias2, m, n, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 FMA kernel with 16-byte global loads: the nn.Linear shapes (forward x W^T, backward-data dY W, weight
// gradient dY^T X with split-K).  Each operand is contiguous either along k or along its tile direction; a thread
// loads ONE float4 per operand and k-tile (instead of four predicated scalars) and the k-contiguous case is
// transposed on its way into shared memory.  Same 64x64x16 tile / 4x4 register tile / epilogue as gemm_kernel.
// Requires: dimensions along the vector direction multiples of 4, 16-byte aligned bases and leading dimensions.
// ---------------------------------------------------------------------------------------------
template <int KFAST>
__device__ __forceinline__ float4 vec_tile_load(const float* __restrict__ X, int64_t ld, int64_t mn0, int64_t k0, int64_t MN,
                                                int64_t K, int tid) {
  // KFAST: X[(mn) * ld + k], thread -> (row = tid / 4, kq = tid % 4);  else X[k * ld + mn], thread -> (k = tid / 16, q = tid % 16)
  int64_t mn, k;
  if (KFAST) { mn = mn0 + (tid >> 2); k = k0 + 4 * (tid & 3); }
  else { k = k0 + (tid >> 4); mn = mn0 + 4 * (tid & 15); }
  if (mn >= MN || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
  return __ldg(reinterpret_cast<const float4*>(X + (KFAST ? mn * ld + k : k * ld + mn)));
}
template <int KFAST>
__device__ __forceinline__ void vec_tile_store(float (*S)[BM + PAD], const float4& v, int tid) {
  if (KFAST) {
    const int row = tid >> 2, kq = 4 * (tid & 3);
    S[kq][row] = v.x; S[kq + 1][row] = v.y; S[kq + 2][row] = v.z; S[kq + 3][row] = v.w;
  } else {
    *reinterpret_cast<float4*>(&S[tid >> 4][4 * (tid & 15)]) = v;
  }
}

template <int A_KF, int B_KF>
__global__ void __launch_bounds__(256, 3) gemm_fp32_vec_kernel(GemmP p) {
  static_assert(BM == BN, "shared tile helper assumes square tiles");
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
  const int64_t lda = A_KF ? p.sam : p.sak, ldb = B_KF ? p.sbn : p.sbk;

  float acc[4][4], rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tx = tid % 16, ty = tid / 16;
  float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
  if (kt_begin < kt_end) {
    ra = vec_tile_load<A_KF>(A, lda, m0, kt_begin * BK, p.M, p.K, tid);
    rb = vec_tile_load<B_KF>(B, ldb, n0, kt_begin * BK, p.N, p.K, tid);
  }
  for (int64_t kt = kt_begin; kt < kt_end; ++kt) {
    vec_tile_store<A_KF>(As, ra, tid);
    vec_tile_store<B_KF>(Bs, rb, tid);
    __syncthreads();
    if (kt + 1 < kt_end) {
      ra = vec_tile_load<A_KF>(A, lda, m0, (kt + 1) * BK, p.M, p.K, tid);
      rb = vec_tile_load<B_KF>(B, ldb, n0, (kt + 1) * BK, p.N, p.K, tid);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rs[i] += av[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  const int64_t pstride = p.M * p.N + (p.rowsum ? p.M : 0);
  float* P = p.splitk > 1 ? p.partial + (int64_t)blockIdx.z * pstride : nullptr;
  if (p.rowsum && blockIdx.y == 0 && tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + ty * 4 + i;
      if (m < p.M) {
        if (P) P[p.M * p.N + m] = rs[i];
        else p.rowsum[m] += rs[i];
      }
    }
  }
  float* C = p.C + (int64_t)batch * p.bsC;
  const float* bias = p.bias ? p.bias + (int64_t)batch * p.bsBias : nullptr;
  const float* bias2 = p.bias2 ? p.bias2 + (int64_t)batch * p.bsBias : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      if (P) P[m * p.N + n] = acc[i][j];
      else store_out(p, C, bias, bias2, m, n, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core kernel (mma.sync.m16n8k8 TF32, fp32 accumulate): 128x64x16 block tile, 8 warps as 4(m) x 2(n),
// warp tile 32x32.  X3 = error-compensated "3xTF32" (a_hi b_hi + a_hi b_lo + a_lo b_hi), fp32-grade accuracy
// for the conv stack whose gradients do not tolerate plain TF32 (measured 6e-3 rel-L2 vs the 1e-3 budget).
// These contractions are HBM-bound (K <= 320), which is why the warp-level MMA path suffices here; the
// on-chip-resident recurrent kernel uses tcgen05 + TMEM (lstm_tc.cu).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CONV, int X3>
__global__ void __launch_bounds__(256, 2) gemm_mma_kernel(GemmP p) {
  __shared__ __align__(16) float As[TBK][TBM + TPAD];
  __shared__ __align__(16) float Bs[TBK][TBN + TPAD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int batch = blockIdx.z / p.splitk, split = blockIdx.z % p.splitk;
  const float* __restrict__ A = p.A + (int64_t)batch * p.bsA;
  const float* __restrict__ B = p.B + (int64_t)batch * p.bsB;
  const int64_t m0 = (int64_t)blockIdx.x * TBM;
  const int64_t n0 = (int64_t)blockIdx.y * TBN;
  const int64_t ktiles = (p.K + TBK - 1) / TBK;
  const int64_t per = (ktiles + p.splitk - 1) / p.splitk;
  const int64_t kt_begin = (int64_t)split * per;
  const int64_t kt_end = min(ktiles, kt_begin + per);

  ALoader<8, TBM, TBK, CONV> la;
  BLoader<4, TBN, TBK, CONV> lb;
  la.init(p, m0, kt_begin * TBK, tid);
  lb.init(p, n0, kt_begin * TBK, tid);

  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  const int g = lane >> 2, q = lane & 3;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

  float ra[8], rb[4];
  if (kt_begin < kt_end) {
    la.load(p, A, m0, kt_begin * TBK, ra);
    lb.load(p, B, n0, kt_begin * TBK, rb);
  }
  for (int64_t kt = kt_begin; kt < kt_end; ++kt) {
#pragma unroll
    for (int e = 0; e < 8; ++e) As[la.kk(e)][la.mm(e)] = ra[e];
#pragma unroll
    for (int e = 0; e < 4; ++e) Bs[lb.kk(e)][lb.nn(e)] = rb[e];
    __syncthreads();
    if (kt + 1 < kt_end) {
      la.load(p, A, m0, (kt + 1) * TBK, ra);
      lb.load(p, B, n0, (kt + 1) * TBK, rb);
    }
#pragma unroll
    for (int ks = 0; ks < TBK; ks += 8) {
      float af[2][4], bf[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int mb = wm + i * 16;
        af[i][0] = As[ks + q][mb + g];
        af[i][1] = As[ks + q][mb + g + 8];
        af[i][2] = As[ks + q + 4][mb + g];
        af[i][3] = As[ks + q + 4][mb + g + 8];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nb = wn + j * 8;
        bf[j][0] = Bs[ks + q][nb + g];
        bf[j][1] = Bs[ks + q + 4][nb + g];
      }
      uint32_t ah[2][4], bh[4][2], al[2][4], bl[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ah[i][c] = to_tf32(af[i][c]);
          if (X3) al[i][c] = to_tf32(af[i][c] - __uint_as_float(ah[i][c]));
        }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          bh[j][c] = to_tf32(bf[j][c]);
          if (X3) bl[j][c] = to_tf32(bf[j][c] - __uint_as_float(bh[j][c]));
        }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (X3) {
            mma_tf32(acc[i][j], al[i], bh[j]);
            mma_tf32(acc[i][j], ah[i], bl[j]);
          }
          mma_tf32(acc[i][j], ah[i], bh[j]);
        }
    }
    __syncthreads();
  }

  // epilogue: thread holds C[(g | g+8)][2q, 2q+1] of each 16x8 tile
  float* C = p.C + (int64_t)batch * p.bsC;
  float* P = p.splitk > 1 ? p.partial + (int64_t)blockIdx.z * p.M * p.N : nullptr;
  const float* bias = p.bias ? p.bias + (int64_t)batch * p.bsBias : nullptr;
  const float* bias2 = p.bias2 ? p.bias2 + (int64_t)batch * p.bsBias : nullptr;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int64_t m = m0 + wm + i * 16 + g + rr * 8;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int64_t n = n0 + wn + j * 8 + 2 * q + cc;
          if (n >= p.N) continue;
          const float v = acc[i][j][rr * 2 + cc];
          if (P) P[m * p.N + n] = v;
          else store_out(p, C, bias, bias2, m, n, v);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Vectorised tensor-core variant for operands that have a unit-stride dimension (all LSTM gradient GEMMs):
// 16-byte cp.async copies straight into the shared-memory layout the fragments are read from (no register
// staging, no transposition), 3-stage pipeline, 128x64x32 tiles.  A_KFAST: A(m,k) is contiguous in k (row-major
// activations) -> tile [m][k]; otherwise contiguous in m (transposed operand of a weight gradient) -> tile [k][m].
// B_NFAST: B(k,n) contiguous in n -> tile [k][n]; otherwise contiguous in k -> tile [n][k].
// Requirements (checked by the launcher): the other stride is a multiple of 4 floats, bases 16-byte aligned,
// M / N / K multiples of 4, no conv window.
// ---------------------------------------------------------------------------------------------
constexpr int VBK = 32, VSTAGES = 3;
constexpr int VA_KF_LD = VBK + 4;    // [128][36]
constexpr int VA_MF_LD = TBM + 8;    // [32][136]
constexpr int VB_NF_LD = TBN + 8;    // [32][72]
constexpr int VB_KF_LD = VBK + 4;    // [64][36]
constexpr int VA_FLOATS = TBM * VA_KF_LD > VBK * VA_MF_LD ? TBM * VA_KF_LD : VBK * VA_MF_LD;  // 4608
constexpr int VB_FLOATS = VBK * VB_NF_LD > TBN * VB_KF_LD ? VBK * VB_NF_LD : TBN * VB_KF_LD;  // 2304
constexpr int VSTAGE_FLOATS = VA_FLOATS + VB_FLOATS;

__device__ __forceinline__ void cp_async16_zfill(float* dst_smem, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

template <int A_KFAST, int B_NFAST>
__global__ void __launch_bounds__(256, 2) gemm_mma_vec_kernel(GemmP p) {
  extern __shared__ __align__(16) float vsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int batch = blockIdx.z / p.splitk, split = blockIdx.z % p.splitk;
  const float* __restrict__ A = p.A + (int64_t)batch * p.bsA;
  const float* __restrict__ B = p.B + (int64_t)batch * p.bsB;
  const int64_t m0 = (int64_t)blockIdx.x * TBM;
  const int64_t n0 = (int64_t)blockIdx.y * TBN;
  const int64_t ktiles = (p.K + VBK - 1) / VBK;
  const int64_t per = (ktiles + p.splitk - 1) / p.splitk;
  const int64_t kt_begin = (int64_t)split * per;
  const int64_t kt_end = min(ktiles, kt_begin + per);
  const int nkt = (int)(kt_end - kt_begin);

  auto issue = [&](int kt_local, int stage) {
    float* sa = vsm + stage * VSTAGE_FLOATS;
    float* sb = sa + VA_FLOATS;
    const int64_t k0 = (kt_begin + kt_local) * VBK;
    // A tile: 128 x 32 floats = 1024 float4, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      if (A_KFAST) {
        const int mm = idx >> 3, k4 = (idx & 7) * 4;
        const int64_t m = m0 + mm, k = k0 + k4;
        const bool v = m < p.M && k < p.K;
        cp_async16_zfill(sa + mm * VA_KF_LD + k4, A + (v ? m * p.sam + k : 0), v);
      } else {
        const int kk = idx >> 5, m4 = (idx & 31) * 4;
        const int64_t m = m0 + m4, k = k0 + kk;
        const bool v = m < p.M && k < p.K;
        cp_async16_zfill(sa + kk * VA_MF_LD + m4, A + (v ? k * p.sak + m : 0), v);
      }
    }
    // B tile: 32 x 64 floats = 512 float4, 2 per thread
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int idx = tid + e * 256;
      if (B_NFAST) {
        const int kk = idx >> 4, n4 = (idx & 15) * 4;
        const int64_t k = k0 + kk, n = n0 + n4;
        const bool v = k < p.K && n < p.N;
        cp_async16_zfill(sb + kk * VB_NF_LD + n4, B + (v ? k * p.sbk + n : 0), v);
      } else {
        const int nn = idx >> 3, k4 = (idx & 7) * 4;
        const int64_t k = k0 + k4, n = n0 + nn;
        const bool v = k < p.K && n < p.N;
        cp_async16_zfill(sb + nn * VB_KF_LD + k4, B + (v ? n * p.sbn + k : 0), v);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  const int g = lane >> 2, q = lane & 3;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

  for (int s = 0; s < VSTAGES - 1; ++s) {
    if (s < nkt) issue(s, s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nkt; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(VSTAGES - 2) : "memory");
    __syncthreads();
    // prefetch tile kt + STAGES - 1 into the stage consumed at iteration kt - 1 (safe after the barrier above)
    if (kt + VSTAGES - 1 < nkt) issue(kt + VSTAGES - 1, (kt + VSTAGES - 1) % VSTAGES);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const float* sa = vsm + (kt % VSTAGES) * VSTAGE_FLOATS;
    const float* sb = sa + VA_FLOATS;
#pragma unroll
    for (int ks = 0; ks < VBK; ks += 8) {
      uint32_t af[2][4], bf[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int mb = wm + i * 16;
        if (A_KFAST) {
          af[i][0] = to_tf32(sa[(mb + g) * VA_KF_LD + ks + q]);
          af[i][1] = to_tf32(sa[(mb + g + 8) * VA_KF_LD + ks + q]);
          af[i][2] = to_tf32(sa[(mb + g) * VA_KF_LD + ks + q + 4]);
          af[i][3] = to_tf32(sa[(mb + g + 8) * VA_KF_LD + ks + q + 4]);
        } else {
          af[i][0] = to_tf32(sa[(ks + q) * VA_MF_LD + mb + g]);
          af[i][1] = to_tf32(sa[(ks + q) * VA_MF_LD + mb + g + 8]);
          af[i][2] = to_tf32(sa[(ks + q + 4) * VA_MF_LD + mb + g]);
          af[i][3] = to_tf32(sa[(ks + q + 4) * VA_MF_LD + mb + g + 8]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nb = wn + j * 8;
        if (B_NFAST) {
          bf[j][0] = to_tf32(sb[(ks + q) * VB_NF_LD + nb + g]);
          bf[j][1] = to_tf32(sb[(ks + q + 4) * VB_NF_LD + nb + g]);
        } else {
          bf[j][0] = to_tf32(sb[(nb + g) * VB_KF_LD + ks + q]);
          bf[j][1] = to_tf32(sb[(nb + g) * VB_KF_LD + ks + q + 4]);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], af[i], bf[j]);
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");

  float* C = p.C + (int64_t)batch * p.bsC;
  float* P = p.splitk > 1 ? p.partial + (int64_t)blockIdx.z * p.M * p.N : nullptr;
  const float* bias = p.bias ? p.bias + (int64_t)batch * p.bsBias : nullptr;
  const float* bias2 = p.bias2 ? p.bias2 + (int64_t)batch * p.bsBias : nullptr;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int64_t m = m0 + wm + i * 16 + g + rr * 8;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int64_t n = n0 + wn + j * 8 + 2 * q + cc;
          if (n >= p.N) continue;
          const float v = acc[i][j][rr * 2 + cc];
          if (P) P[m * p.N + n] = v;
          else store_out(p, C, bias, bias2, m, n, v);
        }
      }
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

// split-K reduction of a GEMM that also carried row sums: partial = [S][MN + M]
__global__ void __launch_bounds__(256) reduce_partials_rs_kernel(const float* __restrict__ P, int S, int64_t MN, int64_t M,
                                                                 float* __restrict__ out, int accumulate,
                                                                 float* __restrict__ rowsum) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = MN + M;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += P[(int64_t)k * n + i];
  if (i < MN) out[i] = accumulate ? out[i] + s : s;
  else rowsum[i - MN] += s;
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
  const bool tf32 = ctx->math_mode >= 1;
  const int64_t tiles = cdiv64(M, tf32 ? TBM : BM) * cdiv64(N, tf32 ? TBN : BN) * nbatch;
  const int64_t ktiles = cdiv64(K, BK);
  int64_t want = cdiv64(2 * (int64_t)ctx->sm_count, tiles);
  int64_t maxs = ktiles / 8;  // at least 8 k-tiles per split
  if (want > maxs) want = maxs;
  if (want > kMaxSplit) want = kMaxSplit;
  if (want < 1) want = 1;
  return (int)want;
}

// + 2048 floats per split: room for the row sums a weight-gradient GEMM may carry (GemmP::rowsum, M <= 2048)
int64_t gemm_splitk_ws_floats(int64_t M, int64_t N, int nbatch) { return (int64_t)kMaxSplit * (M * N + 2048) * nbatch; }

int gemm_launch(wgg_ctx* ctx, const GemmP& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return WGG_OK;
  if (p.M >= (1ll << 31) || p.K >= (1ll << 31)) return wgg_fail(ctx, WGG_EINVAL, "gemm: dimension too large%s");
  if (p.splitk > 1 && (!p.partial || p.act != ACT_NONE || p.bias || p.bias2))
    return wgg_fail(ctx, WGG_EINVAL, "gemm: split-K needs a partial buffer and a plain epilogue%s");
  // large contiguous-K contractions of the tensor-core modes: tcgen05 + TMA engine (gemm_tc.cu)
  if (gemm_tc_usable(ctx, p)) return gemm_tc_launch(ctx, p, st);
  if (p.out_chunk) return wgg_fail(ctx, WGG_EUNSUPPORTED, "gemm: the chunked gate-buffer output exists on the tcgen05 engine only%s");
  const bool tf32 = ctx->math_mode >= 1 && !p.force_fp32 && p.K >= 8 && p.M * p.N >= 4096;
  if (p.rowsum && (tf32 || p.nbatch != 1 || p.conv_mode != 0 || p.M > 2048))
    return wgg_fail(ctx, WGG_EINVAL, "gemm: row sums are only carried by the plain fp32 kernels%s");
  // fp32 with 16-byte loads: every operand contiguous along k or along its tile direction, vector direction % 4
  const bool fa_kf = p.sak == 1, fa_mf = p.sam == 1, fb_kf = p.sbk == 1, fb_nf = p.sbn == 1;
  const bool fvec = !tf32 && p.conv_mode == 0 && (fa_kf || fa_mf) && (fb_kf || fb_nf) &&
                    (fa_kf ? (p.K % 4 == 0 && p.sam % 4 == 0) : (p.M % 4 == 0 && p.sak % 4 == 0)) &&
                    (fb_kf ? (p.K % 4 == 0 && p.sbn % 4 == 0) : (p.N % 4 == 0 && p.sbk % 4 == 0)) &&
                    p.bsA % 4 == 0 && p.bsB % 4 == 0 && (((uintptr_t)p.A | (uintptr_t)p.B) % 16 == 0);
  // small fp32 problems (the Linear layers): halve the M tile when the 64-row tiling would leave SMs idle
  const bool small = !tf32 && !fvec && p.conv_mode == 0 &&
                     cdiv64(p.M, BM) * cdiv64(p.N, BN) * p.nbatch * p.splitk < 2 * (int64_t)ctx->sm_count;
  const int bm = tf32 ? TBM : (small ? 32 : BM), bn = tf32 ? TBN : BN;
  dim3 grid((unsigned)cdiv64(p.M, bm), (unsigned)cdiv64(p.N, bn), (unsigned)(p.nbatch * p.splitk));
  if (grid.y > 65535 || grid.z > 65535) return wgg_fail(ctx, WGG_EINVAL, "gemm: grid too large%s");
  ProfScope prof(ctx, "gemm_kernel", st, 2.0 * (double)p.M * (double)p.N * (double)p.K * p.nbatch,
                 4.0 * ((double)p.M * p.K + (double)p.K * p.N + (double)p.M * p.N) * p.nbatch, p.tag);
  const bool a_kf = p.sak == 1, a_mf = p.sam == 1, b_nf = p.sbn == 1, b_kf = p.sbk == 1;
  const bool vec_ok = tf32 && ctx->math_mode == 1 && p.conv_mode == 0 && !p.x3 && (a_kf || a_mf) && (b_nf || b_kf) &&
                      (p.M % 4 == 0) && (p.N % 4 == 0) && (p.K % 4 == 0) &&
                      ((a_kf ? p.sam : p.sak) % 4 == 0) && ((b_nf ? p.sbk : p.sbn) % 4 == 0) && (p.bsA % 4 == 0) &&
                      (p.bsB % 4 == 0) && (((uintptr_t)p.A | (uintptr_t)p.B) % 16 == 0);
  if (vec_ok) {
    constexpr size_t vsmem = (size_t)VSTAGES * VSTAGE_FLOATS * sizeof(float);
    if (!wgg_smem_ok(ctx, gemm_mma_vec_kernel<1, 1>, vsmem) || !wgg_smem_ok(ctx, gemm_mma_vec_kernel<1, 0>, vsmem) ||
        !wgg_smem_ok(ctx, gemm_mma_vec_kernel<0, 1>, vsmem) || !wgg_smem_ok(ctx, gemm_mma_vec_kernel<0, 0>, vsmem))
      return wgg_fail(ctx, WGG_ECUDA, "gemm_mma_vec_kernel: cannot reserve shared memory%s");
    if (a_kf && b_nf) gemm_mma_vec_kernel<1, 1><<<grid, 256, vsmem, st>>>(p);
    else if (a_kf) gemm_mma_vec_kernel<1, 0><<<grid, 256, vsmem, st>>>(p);
    else if (b_nf) gemm_mma_vec_kernel<0, 1><<<grid, 256, vsmem, st>>>(p);
    else gemm_mma_vec_kernel<0, 0><<<grid, 256, vsmem, st>>>(p);
  } else if (tf32) {
    const bool x3 = ctx->math_mode == 2 && (p.conv_mode != 0 || p.x3);  // compensated convs only in tf32x3 mode
    if (p.conv_mode == 1 && x3) gemm_mma_kernel<1, 1><<<grid, 256, 0, st>>>(p);
    else if (p.conv_mode == 1) gemm_mma_kernel<1, 0><<<grid, 256, 0, st>>>(p);
    else if (p.conv_mode == 2 && x3) gemm_mma_kernel<2, 1><<<grid, 256, 0, st>>>(p);
    else if (p.conv_mode == 2) gemm_mma_kernel<2, 0><<<grid, 256, 0, st>>>(p);
    else if (x3) gemm_mma_kernel<0, 1><<<grid, 256, 0, st>>>(p);
    else gemm_mma_kernel<0, 0><<<grid, 256, 0, st>>>(p);
  } else if (fvec) {
    if (a_kf && b_kf) gemm_fp32_vec_kernel<1, 1><<<grid, 256, 0, st>>>(p);
    else if (a_kf) gemm_fp32_vec_kernel<1, 0><<<grid, 256, 0, st>>>(p);
    else if (b_kf) gemm_fp32_vec_kernel<0, 1><<<grid, 256, 0, st>>>(p);
    else gemm_fp32_vec_kernel<0, 0><<<grid, 256, 0, st>>>(p);
  } else if (p.conv_mode == 1) gemm_kernel<1, 64><<<grid, 256, 0, st>>>(p);
  else if (p.conv_mode == 2) gemm_kernel<2, 64><<<grid, 256, 0, st>>>(p);
  else if (small) gemm_kernel<0, 32><<<grid, 256, 0, st>>>(p);
  else gemm_kernel<0, 64><<<grid, 256, 0, st>>>(p);
  WGG_CHECK_LAUNCH(ctx, "gemm_kernel");
  if (p.splitk > 1) {
    if (p.scn != 1 || p.scm != p.N)
      return wgg_fail(ctx, WGG_EINVAL, "gemm: split-K output must be dense row-major%s");
    if (p.rowsum) {
      const int64_t n = p.M * p.N + p.M;
      reduce_partials_rs_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(p.partial, p.splitk, p.M * p.N, p.M, p.C, p.accumulate,
                                                                           p.rowsum);
      WGG_CHECK_LAUNCH(ctx, "reduce_partials_rs_kernel");
      return WGG_OK;
    }
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

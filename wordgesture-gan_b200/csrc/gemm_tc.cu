// TF32 GEMM on 5th-generation tensor cores for the scaled-model regime (BASELINE configs[3]: gen_hidden_dim 128 ... 1024):
//     C[M, N] (+)= A[M, K] * B[N, K]^T (+ bias + bias2)        A, B row-major with K contiguous ("NT")
// - the shape of every forward contraction of an LSTM layer: the input projection x W_ih^T (M = T * B) and the per-step
// recurrent product h_{t-1} W_hh^T (src/gan/models.py:160).  256 x 256 output tile per CTA (two M = 128 x N = 256
// tcgen05.mma accumulators filling all 512 TMEM columns), K in slabs of 32 floats:
//   * operands arrive by TMA (cp.async.bulk.tensor.2d, one 256-row x 128-byte box per operand and stage, SWIZZLE_128B,
//     out-of-range rows / columns zero-filled by the hardware) into a 3-stage ring, completion on mbarriers;
//   * one elected thread issues the MMAs (K-major SWIZZLE_128B shared-memory descriptors, K advanced by 32 bytes inside
//     the swizzle atom), tcgen05.commit releases the stage;
//   * four epilogue warps read the accumulators with tcgen05.ld and write C rows (bias add or read-modify-write).
// TF32 operands are 4 bytes each, so the tile has to be this large for the tensor pipe not to starve on L2 bandwidth.
// Operand values are taken as they are (the tensor core reads the upper 19 bits of each fp32: truncation).
#include <cuda.h>

#include "tc_common.cuh"

namespace gtc {

using namespace tcu;

constexpr int BM = 256, BN = 256, BK = 32, NST = 3;
constexpr int TILE_BYTES = 256 * BK * 4;  // 32 KB per operand and stage
constexpr int THREADS = 192;              // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue

struct Params {
  CUtensorMap a[2], b[2];  // per batch entry (the two LSTM directions)
  float* c[2];
  const float* bias[2];
  const float* bias2[2];
  int M, N, K, ldc, accumulate;
  int* gerr;
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  // K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor): LBO = 1 (unused), SBO = 1024 B between 8-row groups, version 1
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) gemm_tc_nt_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the ring must sit on a 1024-byte boundary (swizzle atoms are address based)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + NST * TILE_BYTES;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_b + NST * TILE_BYTES);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * NST + 1);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = blockIdx.z;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const uint32_t bar0 = smem_u32(s_bar);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
  const uint32_t DONE = bar0 + 8u * (2 * NST);
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    mbar_init(DONE, 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int KT = (p.K + BK - 1) / BK;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < KT; ++it) {
        const int s = it % NST;
        if (!mbar_wait(EMPTY(s), (uint32_t)(((it / NST) & 1) ^ 1), s_abort, p.gerr, 71)) break;
        mbar_expect_tx(FULL(s), 2 * TILE_BYTES);
        tma_load_2d(smem_u32(s_a + s * TILE_BYTES), &p.a[z], it * BK, m0, FULL(s));
        tma_load_2d(smem_u32(s_b + s * TILE_BYTES), &p.b[z], it * BK, n0, FULL(s));
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, BN);
    bool ok = true;
    for (int it = 0; it < KT && ok; ++it) {
      const int s = it % NST;
      if (!mbar_wait(FULL(s), (uint32_t)((it / NST) & 1), s_abort, p.gerr, 72)) { ok = false; break; }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = smem_u32(s_a + s * TILE_BYTES), b0 = smem_u32(s_b + s * TILE_BYTES);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t bd = desc_sw128(b0 + ks * 32);
          mma_tf32_ss(tmem_base, desc_sw128(a0 + ks * 32), bd, idesc, (it | ks) ? 1u : 0u);
          mma_tf32_ss(tmem_base + 256u, desc_sw128(a0 + 128 * 128 + ks * 32), bd, idesc, (it | ks) ? 1u : 0u);
        }
        mma_commit(EMPTY(s));
      }
      __syncwarp();
    }
    if (ok && elect_one()) mma_commit(DONE);
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    if (mbar_wait(DONE, 0, s_abort, p.gerr, 73)) {
      tc_fence_after();
      float* C = p.c[z];
      const float* bias = p.bias[z];
      const float* bias2 = p.bias2[z];
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int row = m0 + half * 128 + quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 256);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);  // warp-collective: every lane takes part, also for rows beyond M
          if (row < p.M) {
            float* crow = C + (int64_t)row * p.ldc + n0 + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = n0 + c0 + 4 * j;
              if (col < p.N) {
                float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (bias) { const float4 b = *reinterpret_cast<const float4*>(bias + col); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
                if (bias2) { const float4 b = *reinterpret_cast<const float4*>(bias2 + col); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
                if (p.accumulate) { const float4 c = *reinterpret_cast<const float4*>(crow + 4 * j); o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
                *reinterpret_cast<float4*>(crow + 4 * j) = o;
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// [rows, K] fp32 matrix with row stride ld floats -> boxes of 256 rows x 32 floats, 128-byte swizzle, zero fill
bool make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t K, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BK, 256};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gtc

// Can this contraction go through the tcgen05 GEMM?  (fully contiguous-K operands, 16-byte aligned rows, unit column
// stride of C, no activation, no conv window, no split-K)
bool gemm_tc_usable(const wgg_ctx* ctx, const GemmP& p) {
  if (ctx->math_mode < 1 || p.force_fp32 || p.conv_mode || p.splitk > 1 || p.act != ACT_NONE || p.rowsum) return false;
  if (p.sak != 1 || p.sbk != 1 || p.scn != 1 || p.nbatch > 2) return false;
  if ((p.sam & 3) || (p.sbn & 3) || (p.scm & 3) || (p.N & 3) || (p.K & 3)) return false;
  if (p.M < 128 || p.N < 128 || p.K < 32) return false;  // small problems stay on the mma.sync engine
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al(p.A) || !al(p.B) || !al(p.C) || (p.bias && !al(p.bias)) || (p.bias2 && !al(p.bias2))) return false;
  if (p.nbatch == 2 && ((p.bsA & 3) || (p.bsB & 3) || (p.bsC & 3) || (p.bsBias & 3))) return false;
  return gtc::encode_fn() != nullptr;
}

int gemm_tc_launch(wgg_ctx* ctx, const GemmP& p, cudaStream_t st) {
  gtc::Params prm;
  memset(&prm, 0, sizeof(prm));
  for (int z = 0; z < p.nbatch; ++z) {
    if (!gtc::make_map(&prm.a[z], p.A + z * p.bsA, p.M, p.K, p.sam) || !gtc::make_map(&prm.b[z], p.B + z * p.bsB, p.N, p.K, p.sbn))
      return wgg_fail(ctx, WGG_ECUDA, "gemm_tc: cuTensorMapEncodeTiled failed%s");
    prm.c[z] = p.C + z * p.bsC;
    prm.bias[z] = p.bias ? p.bias + z * p.bsBias : nullptr;
    prm.bias2[z] = p.bias2 ? p.bias2 + z * p.bsBias : nullptr;
  }
  prm.M = (int)p.M; prm.N = (int)p.N; prm.K = (int)p.K; prm.ldc = (int)p.scm; prm.accumulate = p.accumulate;
  prm.gerr = ctx->async_err;
  constexpr size_t smem = (size_t)2 * gtc::NST * gtc::TILE_BYTES + (2 * gtc::NST + 1) * 8 + 16 + 1024;
  if (!wgg_smem_ok(ctx, gtc::gemm_tc_nt_kernel, smem)) return wgg_fail(ctx, WGG_ECUDA, "gemm_tc_nt_kernel: cannot reserve shared memory%s");
  dim3 grid((unsigned)cdiv64(p.N, gtc::BN), (unsigned)cdiv64(p.M, gtc::BM), (unsigned)p.nbatch);
  ProfScope prof(ctx, "gemm_tc_nt_kernel", st, 2.0 * p.M * (double)p.N * p.K * p.nbatch,
                 4.0 * p.nbatch * ((double)p.M * p.K + (double)p.N * p.K + (double)p.M * p.N * (p.accumulate ? 2 : 1)),
                 p.tag ? p.tag : "gemm_tc_nt_kernel");
  gtc::gemm_tc_nt_kernel<<<grid, gtc::THREADS, smem, st>>>(prm);
  WGG_CHECK_LAUNCH(ctx, "gemm_tc_nt_kernel");
  return WGG_OK;
}

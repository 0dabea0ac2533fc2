// TF32 GEMM on 5th-generation tensor cores for the scaled-model regime (BASELINE configs[3]: gen_hidden_dim 128 ... 1024):
//     C[M, N] (+)= A[M, K] * B[N, K]^T (+ bias + bias2)        A, B row-major with K contiguous ("NT")
// - the shape of every forward contraction of an LSTM layer: the input projection x W_ih^T (M = T * B) and the per-step
// recurrent product h_{t-1} W_hh^T (src/gan/models.py:160).  256 x 256 output tile per CTA (two M = 128 x N = 256
// tcgen05.mma accumulators filling all 512 TMEM columns), K in slabs of 32 floats:
//   * operands arrive by TMA (cp.async.bulk.tensor.2d, one 256-row x 128-byte box per operand and stage, SWIZZLE_128B,
//     out-of-range rows / columns zero-filled by the hardware) into a 3-stage ring, completion on mbarriers;
//   * one elected thread issues the MMAs (K-major SWIZZLE_128B shared-memory descriptors, K advanced by 32 bytes inside
//     the swizzle atom), tcgen05.commit releases the stage;
//   * eight epilogue warps (two per TMEM lane quarter, one per 128-column half) read the accumulators with tcgen05.ld and
//     write C rows (bias - staged once in shared memory - add, or read-modify-write), or - GemmP::out_chunk - the gate buffer
//     in the chunked order [t][tile of 128 gestures][columns / 4][128 rows][4 floats] of the recurrence kernels;
//   * split-K (the weight gradients: K = T * B): blockIdx.z also enumerates k ranges, each writing its own dense partial
//     slice, reduced in fixed order.  Weight / input gradients reach this "NT" form through K-major TF32 images of their
//     operands (transpose_tf32_kernel; unchunk_da_kernel in lstm.cu).
// TF32 operands are 4 bytes each, so the tile has to be this large for the tensor pipe not to starve on L2 bandwidth.
// Operand values are taken as they are (the tensor core reads the upper 19 bits of each fp32: truncation); the producers of
// h, da and the transposed images round to nearest.
//
// The same pipeline carries the recurrence of the scaled regime with the LSTM cell fused into the epilogue:
//   * gemm_tc_lstm_fwd_kernel / gemm_tc_lstm_bwd_kernel: one launch per timestep and layer (any hidden size), programmatic
//     dependent launch between consecutive steps;
//   * lstm128_tc_fwd_kernel: persistent over all timesteps for gen_hidden_dim = 128 (h tile and cell state stay on the SM).
#include <cuda.h>

#include "tc_common.cuh"

namespace gtc {

using namespace tcu;

constexpr int BM = 256, BN = 256, BK = 32, NST = 3;
constexpr int TILE_BYTES = 256 * BK * 4;  // 32 KB per operand and stage
constexpr int THREADS = 320;              // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue
constexpr int EPI_THREADS = 256;

struct Params {
  CUtensorMap a[2], b[2];  // per batch entry (the two LSTM directions)
  float* c[2];
  const float* bias[2];
  const float* bias2[2];
  int M, N, K, ldc, accumulate;
  // split-K: blockIdx.z = batch entry * splits + split; a split contracts k in [split * k_len, (split + 1) * k_len) (k_len a
  // multiple of BK) and writes its own dense [M][N] slice of the partial buffer (c[z] + split * M * N); splits == 1: plain
  int splits, k_len;
  // out_chunk: C is the gate buffer of the persistent H = 128 recurrence, row m = t * chunk_B + b, written in the order
  // its epilogue reads: [t][b / 128][N / 4][b % 128][4 floats] (lane = gesture: coalesced 16-byte accesses on both sides)
  int out_chunk, chunk_B, chunk_tiles;
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
  static_assert(2 * NST + 1 <= 8, "barrier block is 64 bytes");
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 8);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  float* s_bias = reinterpret_cast<float*>(s_tmem + 4);  // [BN] bias + bias2 of this column tile (16-byte aligned)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
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
  const int k_begin = split * p.k_len;
  const int k_end = k_begin + p.k_len < p.K ? k_begin + p.k_len : p.K;  // the last split's ragged tail is zero-filled by TMA
  const int KT = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < KT; ++it) {
        const int s = it % NST;
        if (!mbar_wait(EMPTY(s), (uint32_t)(((it / NST) & 1) ^ 1), s_abort, p.gerr, 71)) break;
        mbar_expect_tx(FULL(s), 2 * TILE_BYTES);
        tma_load_2d(smem_u32(s_a + s * TILE_BYTES), &p.a[z], k_begin + it * BK, m0, FULL(s));
        tma_load_2d(smem_u32(s_b + s * TILE_BYTES), &p.b[z], k_begin + it * BK, n0, FULL(s));
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
    if (ok && KT > 0 && elect_one()) mma_commit(DONE);
    __syncwarp();
  } else {
    // a warp may read the TMEM lanes 32 * (warp id % 4) ... + 31; the second group of four warps takes the upper column half
    const int quarter = warp & 3, chalf = (warp - 2) >> 2;
    {
      const int et = tid - 64, col = n0 + et;
      float b = 0.f;
      if (col < p.N) {
        if (p.bias[z]) b += __ldg(p.bias[z] + col);
        if (p.bias2[z]) b += __ldg(p.bias2[z] + col);
      }
      s_bias[et] = b;
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
    }
    if (KT == 0 || mbar_wait(DONE, 0, s_abort, p.gerr, 73)) {
      tc_fence_after();
      float* C = p.c[z] + (int64_t)split * p.M * p.N;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int row = m0 + half * 128 + quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 256 + chalf * 128);
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          const int colb = chalf * 128 + c0;  // column inside the tile
          float* crow = C + (int64_t)row * p.ldc + n0 + colb;
          if (p.out_chunk) {  // plain store, bias added; row -> (t, tile, row in tile)
            float v[32];
            tmem_ld32(taddr + c0, v);
            if (row < p.M) {
              const int t = row / p.chunk_B, b = row - t * p.chunk_B;
              float4* dst = reinterpret_cast<float4*>(C) +
                            (((int64_t)t * p.chunk_tiles + (b >> 7)) * (p.N >> 2) + ((n0 + colb) >> 2)) * 128 + (b & 127);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (n0 + colb + 4 * j < p.N) {
                  const float4 bb = *reinterpret_cast<const float4*>(s_bias + colb + 4 * j);
                  dst[(int64_t)j * 128] = make_float4(v[4 * j] + bb.x, v[4 * j + 1] + bb.y, v[4 * j + 2] + bb.z, v[4 * j + 3] + bb.w);
                }
              }
            }
            continue;
          }
          float4 old[8];
          if (p.accumulate && row < p.M) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              old[j] = (n0 + colb + 4 * j < p.N) ? *reinterpret_cast<const float4*>(crow + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float v[32];
          tmem_ld32(taddr + c0, v);  // warp-collective: every lane takes part, also for rows beyond M
          if (KT == 0) {             // an empty split (cannot happen with the launcher's split choice) contributes zeros
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (row < p.M) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (n0 + colb + 4 * j < p.N) {
                const float4 b = *reinterpret_cast<const float4*>(s_bias + colb + 4 * j);
                float4 o = make_float4(v[4 * j] + b.x, v[4 * j + 1] + b.y, v[4 * j + 2] + b.z, v[4 * j + 3] + b.w);
                if (p.accumulate) { o.x += old[j].x; o.y += old[j].y; o.z += old[j].z; o.w += old[j].w; }
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


// ---------------------------------------------------------------------------------------------
// One timestep of the recurrence of a scaled BiLSTM layer (H > 64: W_hh does not fit one SM), both directions:
//     pre = gates[d][t] (input projection + biases)  +  h[d][t_prev] W_hh[d]^T          (tcgen05, TF32)
//     i, f, o = sigmoid(.), g = tanh(.);  c = f c_prev + i g;  h = o tanh(c)           (epilogue, fp32)
// (torch.nn.LSTM as called at src/gan/models.py:160).  CTA = 128 gestures x 64 hidden units x one direction: the B
// operand is the four 64-row slices (one per gate) of W_hh for these units - four TMA boxes stacked into one 256-row
// K-major tile - so the accumulator (256 TMEM columns) holds all four gates of a unit under the same lane and the cell
// update needs no exchange.  A = h[t_prev] through a 3-D tensor map over hseq (k, gesture, t).  The eight epilogue warps
// (lane quarter x 32-unit half) read the input projection / c_prev with 16-byte loads issued BEFORE the wait for the
// tensor core, add the accumulator, and write h (rounded to TF32, as the persistent H = 48 kernel does: it is the next
// step's and the next layer's MMA operand), c and - when the pass carries gradients - the activated gates in place.
// Step 0 has no recurrent term: K = 0, no MMA, the accumulator counts as zero.
// ---------------------------------------------------------------------------------------------
constexpr int SBM = 128, SUN = 64;  // ring depth SNST is a template parameter (2, 3 or 4 stages of 48 KB)
constexpr int SA_BYTES = SBM * BK * 4;      // 16 KB
constexpr int SB_BYTES = 4 * SUN * BK * 4;  // 32 KB
constexpr int SBOX_BYTES = SUN * BK * 4;    // one gate's 64 rows

struct StepFwdParams {
  CUtensorMap a[2];  // per direction: h of that direction inside hseq, dims (H, B, T), strides (2H, B * 2H) floats
  CUtensorMap b[2];  // per direction: W_hh [4H][H], boxes of 64 rows
  float* gates;      // [2][T][B][4H]
  float* hseq;       // [T][B][2H]
  float* cseq;       // store: [2][T][B][H]
  float* cstate;     // no store: [2][B][H]
  int T, B, H, step, store;
  int* gerr;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

// Programmatic dependent launch: consecutive timestep kernels are launched with the programmatic-serialisation attribute,
// so a step's CTAs may start (barrier init, TMEM allocation, the weight slabs of the first ring stages, the input
// projection) while the previous step's epilogue is still running; everything the previous step produces is touched only
// after pdl_wait() (which returns once the previous grid has completed and its writes are visible).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// CHUNK = 1: gate buffer, c (sequence or carried state) in the chunked order [..][tile][columns / 4][128 rows][4 floats] - the
// epilogue's 16-byte accesses are then coalesced (lane = gesture) instead of one 128-byte line per lane.
// CL = 2: the two CTAs of a cluster work on neighbouring row tiles of the same (unit block, direction) and therefore need the
// same weight slabs: each fetches half of every slab (two of the four gate boxes) and multicasts it into both CTAs' rings, so a
// CTA streams A + B / 2 instead of A + B through L2 (the kernel is bound by that stream: 48 KB per k-slab and CTA).  A ring
// stage is refilled only after BOTH CTAs' MMAs have released it (tcgen05.commit multicast onto both EMPTY barriers).
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int SNST, int CHUNK, int CL>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_lstm_fwd_kernel(const __grid_constant__ StepFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + SNST * SA_BYTES;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_b + SNST * SB_BYTES);
  static_assert(2 * SNST + 1 <= 16, "barrier block is 128 bytes");
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 16);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.z;
  const int u0 = blockIdx.x * SUN, m0 = blockIdx.y * SBM;
  const int t = dir ? p.T - 1 - p.step : p.step;
  const int tp = dir ? t + 1 : t - 1;
  const int H = p.H;
  const uint32_t bar0 = smem_u32(s_bar);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (SNST + s); };
  const uint32_t DONE = bar0 + 8u * (2 * SNST);
  if (tid == 0) {
    for (int s = 0; s < SNST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), CL); }
    mbar_init(DONE, 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t crank = 0;
  if (CL > 1) {
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    cluster_sync_all();  // the peer's barriers exist before anything of ours can reach them
  }
  const uint32_t tmem_base = *s_tmem;
  const int KT = p.step > 0 ? (H + BK - 1) / BK : 0;
  // weight boxes of one ring stage: alone, all four gates; in a pair, gates {2 rank, 2 rank + 1} multicast to both CTAs
  auto load_b = [&](int s, int it) {
    if (CL > 1) {
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        const int g = 2 * (int)crank + gg;
        tma_load_2d_mc(smem_u32(s_b + s * SB_BYTES + g * SBOX_BYTES), &p.b[dir], it * BK, g * H + u0, FULL(s), (uint16_t)0x3);
      }
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        tma_load_2d(smem_u32(s_b + s * SB_BYTES + g * SBOX_BYTES), &p.b[dir], it * BK, g * H + u0, FULL(s));
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      // weight slabs of the first ring stages do not depend on the previous timestep: request them, then wait for it
      const int npre = KT < SNST ? KT : SNST;
      for (int it = 0; it < npre; ++it) {
        mbar_expect_tx(FULL(it), SA_BYTES + SB_BYTES);
        load_b(it, it);
      }
      pdl_wait();
      pdl_launch_dependents();
      for (int it = 0; it < npre; ++it) tma_load_3d(smem_u32(s_a + it * SA_BYTES), &p.a[dir], it * BK, m0, tp, FULL(it));
      for (int it = npre; it < KT; ++it) {
        const int s = it % SNST;
        if (!mbar_wait(EMPTY(s), (uint32_t)(((it / SNST) & 1) ^ 1), s_abort, p.gerr, 74)) break;
        mbar_expect_tx(FULL(s), SA_BYTES + SB_BYTES);
        tma_load_3d(smem_u32(s_a + s * SA_BYTES), &p.a[dir], it * BK, m0, tp, FULL(s));
        load_b(s, it);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, 256);
    bool ok = true;
    for (int it = 0; it < KT && ok; ++it) {
      const int s = it % SNST;
      if (!mbar_wait(FULL(s), (uint32_t)((it / SNST) & 1), s_abort, p.gerr, 75)) { ok = false; break; }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = smem_u32(s_a + s * SA_BYTES), b0 = smem_u32(s_b + s * SB_BYTES);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks)
          mma_tf32_ss(tmem_base, desc_sw128(a0 + ks * 32), desc_sw128(b0 + ks * 32), idesc, (it | ks) ? 1u : 0u);
        if (CL > 1) mma_commit_mc(EMPTY(s), (uint16_t)0x3);
        else mma_commit(EMPTY(s));
      }
      __syncwarp();
    }
    if (ok && KT > 0 && elect_one()) mma_commit(DONE);
    __syncwarp();
  } else {
    const int quarter = warp & 3, uh = (warp - 2) >> 2;
    const int row = m0 + quarter * 32 + lane;
    const bool rok = row < p.B;
    const int64_t TB = (int64_t)p.T * p.B;
    const int rl = quarter * 32 + lane;
    const int tiles = (p.B + SBM - 1) / SBM;
    const int64_t rr0 = rok ? row : 0;
    // CHUNK: the 4 columns [n, n + 4) of this row inside a [columns / 4][128][4] block: block + (n >> 2) * 512 + rl * 4
    auto at = [&](float* base, int n) -> float* { return CHUNK ? base + (int64_t)(n >> 2) * 512 : base + n; };
    float* gp = CHUNK ? p.gates + ((((int64_t)dir * p.T + t) * tiles + blockIdx.y) * (int64_t)(4 * H) * SBM + rl * 4)
                      : p.gates + (((int64_t)dir * p.T + t) * p.B + rr0) * 4 * H;
    float* hp = p.hseq + ((int64_t)t * p.B + rr0) * 2 * H + dir * H;
    auto cptr = [&](int tt) -> float* {  // c of this row: stored sequence (tt = timestep) or the carried state
      if (p.store)
        return CHUNK ? p.cseq + ((((int64_t)dir * p.T + tt) * tiles + blockIdx.y) * (H / 4) * SBM + rl) * 4
                     : p.cseq + (int64_t)dir * TB * H + ((int64_t)tt * p.B + rr0) * H;
      return CHUNK ? p.cstate + (((int64_t)dir * tiles + blockIdx.y) * (H / 4) * SBM + rl) * 4 : p.cstate + ((int64_t)dir * p.B + rr0) * H;
    };
    float* cprev = p.step > 0 ? cptr(tp) : nullptr;
    float* cout = cptr(t);
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // four chunks of 8 hidden units per thread (two adjacent 16-byte accesses = one full 32-byte sector per row and gate); a
    // chunk's operands (the four pre-activations, c_prev: ten 16-byte loads) are requested two chunks ahead of their use,
    // the first two before the wait for the tensor core.  The input projection does not depend on the previous timestep
    // and is requested before pdl_wait(), c_prev after it.
    float4 pre[2][4][2], cp[2][2];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_pre = [&](int c, int slot) {
      const int u = u0 + uh * 32 + c * 8;
      const bool ok = rok && u < H;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        pre[slot][g][0] = ok ? ld4(at(gp, g * H + u)) : z4;
        pre[slot][g][1] = ok ? ld4(at(gp, g * H + u + 4)) : z4;
      }
    };
    auto load_c = [&](int c, int slot) {
      const int u = u0 + uh * 32 + c * 8;
      const bool ok = rok && u < H && cprev;
      cp[slot][0] = ok ? ld4(at(cprev, u)) : z4;
      cp[slot][1] = ok ? ld4(at(cprev, u + 4)) : z4;
    };
    load_pre(0, 0);
    load_pre(1, 1);
    pdl_wait();
    load_c(0, 0);
    load_c(1, 1);
    bool live = true;
    if (KT > 0) {
      live = mbar_wait(DONE, 0, s_abort, p.gerr, 76);
      tc_fence_after();
    }
    if (live) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int slot = c & 1;
        const int u = u0 + uh * 32 + c * 8;
        float a[4][8];
        if (KT > 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t rr[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7])
                         : "r"(taddr + (uint32_t)(g * SUN + uh * 32 + c * 8)));
#pragma unroll
            for (int i = 0; i < 8; ++i) a[g][i] = __uint_as_float(rr[i]);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[g][i] = 0.f;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 lo = pre[slot][g][0], hi = pre[slot][g][1];
          a[g][0] += lo.x; a[g][1] += lo.y; a[g][2] += lo.z; a[g][3] += lo.w;
          a[g][4] += hi.x; a[g][5] += hi.y; a[g][6] += hi.z; a[g][7] += hi.w;
        }
        const float cpv[8] = {cp[slot][0].x, cp[slot][0].y, cp[slot][0].z, cp[slot][0].w,
                              cp[slot][1].x, cp[slot][1].y, cp[slot][1].z, cp[slot][1].w};
        if (c + 2 < 4) { load_pre(c + 2, slot); load_c(c + 2, slot); }
        if (rok && u < H) {
          float hv[8], cv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float ig, fg, gg, og, hh;
            lstm_cell_fast(a[0][i], a[1][i], a[2][i], a[3][i], cpv[i], ig, fg, gg, og, cv[i], hh);
            hv[i] = rna_tf32(hh);
            a[0][i] = ig; a[1][i] = fg; a[2][i] = gg; a[3][i] = og;
          }
          st4(hp + u, make_float4(hv[0], hv[1], hv[2], hv[3]));
          st4(hp + u + 4, make_float4(hv[4], hv[5], hv[6], hv[7]));
          st4(at(cout, u), make_float4(cv[0], cv[1], cv[2], cv[3]));
          st4(at(cout, u + 4), make_float4(cv[4], cv[5], cv[6], cv[7]));
          if (p.store) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              st4(at(gp, g * H + u), make_float4(a[g][0], a[g][1], a[g][2], a[g][3]));
              st4(at(gp, g * H + u + 4), make_float4(a[g][4], a[g][5], a[g][6], a[g][7]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------
// One timestep of back-propagation through time of the same layer (reverse scan, both directions):
//     dh_rec = da[d][t_next] W_hh[d]                      (tcgen05; B operand = the transposed image W_hh^T [H][4H])
//     dh = dh_out[t] + dh_rec;  do = dh tanh(c);  dc = dc_carry + dh o (1 - tanh^2 c)
//     da_i = dc g i (1 - i), da_f = dc c_prev f (1 - f), da_g = dc i (1 - g^2), da_o = do o (1 - o);  dc_carry = dc f
// CTA = 128 gestures x 64 hidden units x direction; the epilogue overwrites the activated gates of step t with da (the
// next launch's A operand and the operand of the weight-gradient GEMMs).  The first launch (step T - 1) has K = 0.
// ---------------------------------------------------------------------------------------------
constexpr int WNST = 4;  // (7 stages measured slower: 19.1 vs 17.2 ms of BPTT per H = 128 step)
constexpr int WB_BYTES = SUN * BK * 4;  // 8 KB

struct StepBwdParams {
  CUtensorMap a[2];  // per direction: gates[d] as (4H, B, T)
  CUtensorMap b[2];  // per direction: W_hh^T [H][4H], boxes of 64 rows
  float* gates;
  const float* cseq;
  const float* dh_out;  // [T][B][2H]
  float* dcs;           // [2][B][H] carried dc
  int T, B, H, step;
  int* gerr;
};

// CHUNK = 1 (H = 128 with the chunked stash of the persistent forward): gates / c / the carried dc are
// [..][tile][columns / 4][128 rows][4 floats]; a k-slab (32 columns) of da is then one contiguous 16 KB block in UMMA
// core-matrix order (no swizzle: k-chunks 2 KB apart, 8-row groups 128 B apart) and arrives by one plain bulk copy.
template <int CHUNK>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_lstm_bwd_kernel(const __grid_constant__ StepBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + WNST * SA_BYTES;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_b + WNST * WB_BYTES);
  static_assert(2 * WNST + 1 <= 16, "barrier block is 128 bytes");
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 16);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.z;
  const int u0 = blockIdx.x * SUN, m0 = blockIdx.y * SBM;
  const int t = dir ? p.T - 1 - p.step : p.step;
  const int tn = dir ? t - 1 : t + 1;  // the step processed by the previous launch (later in the sequence's own order)
  const int tp = dir ? t + 1 : t - 1;  // the step whose cell state is c_prev
  const int H = p.H;
  const int tiles = (p.B + SBM - 1) / SBM;
  // chunked gate block of this tile at sequence position tt: [4H / 4 chunks][128 rows][4 floats], 256 KB
  auto gate_blk = [&](int tt) -> float* { return p.gates + (((int64_t)dir * p.T + tt) * tiles + blockIdx.y) * (int64_t)(4 * H) * SBM; };
  const uint32_t bar0 = smem_u32(s_bar);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (WNST + s); };
  const uint32_t DONE = bar0 + 8u * (2 * WNST);
  if (tid == 0) {
    for (int s = 0; s < WNST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    mbar_init(DONE, 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const bool first = p.step == p.T - 1;
  const int KT = first ? 0 : (4 * H + BK - 1) / BK;

  if (warp == 0) {
    if (lane == 0) {
      const int npre = KT < WNST ? KT : WNST;
      for (int it = 0; it < npre; ++it) {
        mbar_expect_tx(FULL(it), SA_BYTES + WB_BYTES);
        tma_load_2d(smem_u32(s_b + it * WB_BYTES), &p.b[dir], it * BK, u0, FULL(it));
      }
      pdl_wait();
      pdl_launch_dependents();
      const uint8_t* ablk = CHUNK ? reinterpret_cast<const uint8_t*>(gate_blk(tn)) : nullptr;
      for (int it = 0; it < npre; ++it) {
        if (CHUNK) bulk_g2s(smem_u32(s_a + it * SA_BYTES), ablk + (int64_t)it * SA_BYTES, SA_BYTES, FULL(it));
        else tma_load_3d(smem_u32(s_a + it * SA_BYTES), &p.a[dir], it * BK, m0, tn, FULL(it));
      }
      for (int it = npre; it < KT; ++it) {
        const int s = it % WNST;
        if (!mbar_wait(EMPTY(s), (uint32_t)(((it / WNST) & 1) ^ 1), s_abort, p.gerr, 77)) break;
        mbar_expect_tx(FULL(s), SA_BYTES + WB_BYTES);
        if (CHUNK) bulk_g2s(smem_u32(s_a + s * SA_BYTES), ablk + (int64_t)it * SA_BYTES, SA_BYTES, FULL(s));
        else tma_load_3d(smem_u32(s_a + s * SA_BYTES), &p.a[dir], it * BK, m0, tn, FULL(s));
        tma_load_2d(smem_u32(s_b + s * WB_BYTES), &p.b[dir], it * BK, u0, FULL(s));
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, SUN);
    bool ok = true;
    for (int it = 0; it < KT && ok; ++it) {
      const int s = it % WNST;
      if (!mbar_wait(FULL(s), (uint32_t)((it / WNST) & 1), s_abort, p.gerr, 78)) { ok = false; break; }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = smem_u32(s_a + s * SA_BYTES), b0 = smem_u32(s_b + s * WB_BYTES);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks)
          mma_tf32_ss(tmem_base, CHUNK ? make_desc(a0 + ks * 4096, 2048, 128) : desc_sw128(a0 + ks * 32), desc_sw128(b0 + ks * 32),
                      idesc, (it | ks) ? 1u : 0u);
        mma_commit(EMPTY(s));
      }
      __syncwarp();
    }
    if (ok && KT > 0 && elect_one()) mma_commit(DONE);
    __syncwarp();
  } else {
    const int quarter = warp & 3, uh = (warp - 2) >> 2;
    const int row = m0 + quarter * 32 + lane;
    const bool rok = row < p.B;
    const int64_t TB = (int64_t)p.T * p.B;
    const int64_t r = rok ? row : 0;
    const int rl = quarter * 32 + lane;
    // CHUNK: element (column n, this row) of a [columns / 4][128][4] block sits at block + (n >> 2) * 512 + rl * 4 + (n & 3)
    float* gp = CHUNK ? gate_blk(t) + rl * 4 : p.gates + (((int64_t)dir * p.T + t) * p.B + r) * 4 * H;
    auto cblk = [&](int tt) -> const float* {
      return p.cseq + ((((int64_t)dir * p.T + tt) * tiles + blockIdx.y) * (H / 4) * SBM + rl) * 4;
    };
    const float* cb = p.cseq + (int64_t)dir * TB * H;
    const float* cc = CHUNK ? cblk(t) : cb + ((int64_t)t * p.B + r) * H;
    const float* cpp = p.step > 0 ? (CHUNK ? cblk(tp) : cb + ((int64_t)tp * p.B + r) * H) : nullptr;
    const float* dho = p.dh_out + ((int64_t)t * p.B + r) * 2 * H + dir * H;
    float* dcs = CHUNK ? p.dcs + ((((int64_t)dir * tiles + blockIdx.y) * (H / 4)) * SBM + rl) * 4 : p.dcs + ((int64_t)dir * p.B + r) * H;
    // address of the 4 columns [n, n + 4) of this row (n a multiple of 4)
    auto at = [&](const float* base, int n) -> const float* { return CHUNK ? base + (int64_t)(n >> 2) * 512 : base + n; };
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // four chunks of 8 hidden units per thread (two adjacent 16-byte accesses = one full 32-byte sector per row and
    // tensor), operands requested one chunk ahead; everything but the carried dc is independent of the previous launch and
    // is requested before pdl_wait()
    struct Ops { float4 gi[2], gf[2], gg[2], go[2], c4[2], cp4[2], dh4[2]; };
    struct Dc { float4 v[2]; };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_ops = [&](int c) {
      const int u = u0 + uh * 32 + c * 8;
      const bool ok = rok && u < H;
      Ops o;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int uq = u + 4 * q;
        o.gi[q] = ok ? ld4(at(gp, uq)) : z4; o.gf[q] = ok ? ld4(at(gp, H + uq)) : z4; o.gg[q] = ok ? ld4(at(gp, 2 * H + uq)) : z4;
        o.go[q] = ok ? ld4(at(gp, 3 * H + uq)) : z4;
        o.c4[q] = ok ? ld4(at(cc, uq)) : z4; o.cp4[q] = (ok && cpp) ? ld4(at(cpp, uq)) : z4; o.dh4[q] = ok ? ld4(dho + uq) : z4;
      }
      return o;
    };
    auto load_dc = [&](int c) {
      const int u = u0 + uh * 32 + c * 8;
      const bool ok = rok && u < H && !first;
      Dc d;
      d.v[0] = ok ? ld4(at(dcs, u)) : z4;
      d.v[1] = ok ? ld4(at(dcs, u + 4)) : z4;
      return d;
    };
    Ops cur = load_ops(0);
    pdl_wait();
    Dc dc_cur = load_dc(0);
    bool live = true;
    if (KT > 0) {
      live = mbar_wait(DONE, 0, s_abort, p.gerr, 79);
      tc_fence_after();
    }
    if (live) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int u = u0 + uh * 32 + c * 8;
        const bool ok = rok && u < H;
        Ops nxt = cur;
        Dc dc_nxt = dc_cur;
        if (c + 1 < 4) { nxt = load_ops(c + 1); dc_nxt = load_dc(c + 1); }
        float rec[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (KT > 0) tmem_ld8(taddr + (uint32_t)(uh * 32 + c * 8), rec);
        if (ok) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float iv[4] = {cur.gi[q].x, cur.gi[q].y, cur.gi[q].z, cur.gi[q].w}, fv[4] = {cur.gf[q].x, cur.gf[q].y, cur.gf[q].z, cur.gf[q].w},
                        gv[4] = {cur.gg[q].x, cur.gg[q].y, cur.gg[q].z, cur.gg[q].w}, ov[4] = {cur.go[q].x, cur.go[q].y, cur.go[q].z, cur.go[q].w},
                        cv[4] = {cur.c4[q].x, cur.c4[q].y, cur.c4[q].z, cur.c4[q].w}, pv[4] = {cur.cp4[q].x, cur.cp4[q].y, cur.cp4[q].z, cur.cp4[q].w},
                        dv[4] = {cur.dh4[q].x, cur.dh4[q].y, cur.dh4[q].z, cur.dh4[q].w},
                        kv[4] = {dc_cur.v[q].x, dc_cur.v[q].y, dc_cur.v[q].z, dc_cur.v[q].w};
            float di[4], df[4], dg[4], dO[4], dk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float tc = tanh_ex2(cv[i]);
              const float dh = dv[i] + rec[4 * q + i];
              const float d_o = dh * tc;
              const float dct = kv[i] + dh * ov[i] * (1.f - tc * tc);
              // da is an MMA operand three times over (next launch, dx, dW): round to nearest instead of letting the
              // tensor core truncate
              di[i] = rna_tf32(dct * gv[i] * iv[i] * (1.f - iv[i]));
              df[i] = rna_tf32(dct * pv[i] * fv[i] * (1.f - fv[i]));
              dg[i] = rna_tf32(dct * iv[i] * (1.f - gv[i] * gv[i]));
              dO[i] = rna_tf32(d_o * ov[i] * (1.f - ov[i]));
              dk[i] = dct * fv[i];
            }
            const int uq = u + 4 * q;
            st4(const_cast<float*>(at(gp, uq)), make_float4(di[0], di[1], di[2], di[3]));
            st4(const_cast<float*>(at(gp, H + uq)), make_float4(df[0], df[1], df[2], df[3]));
            st4(const_cast<float*>(at(gp, 2 * H + uq)), make_float4(dg[0], dg[1], dg[2], dg[3]));
            st4(const_cast<float*>(at(gp, 3 * H + uq)), make_float4(dO[0], dO[1], dO[2], dO[3]));
            st4(const_cast<float*>(at(dcs, uq)), make_float4(dk[0], dk[1], dk[2], dk[3]));
          }
        }
        cur = nxt;
        dc_cur = dc_nxt;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent recurrence for gen_hidden_dim = 128 (first rung of BASELINE configs[3]; torch.nn.LSTM as called at
// src/gan/models.py:160): ONE launch per layer instead of one per timestep.  CTA = 128 gestures x one direction x all T
// steps; the four gates of all 128 units are 512 accumulator columns = the whole TMEM of the SM, as two halves (units
// 0..63 / 64..127, each N = 256) so that the cell math of the first half overlaps the tensor-core work of the second.
//   * h_{t-1} never leaves the SM: the epilogue writes h_t (TF32-rounded) straight into the K-major SWIZZLE_128B operand
//     tile of the next step (two tiles, ping-pong: the second half's MMAs of step t still read h_{t-1} while the first
//     half's epilogue already produces h_t), besides the copy in hseq that the next layer reads;
//   * c stays in registers (64 per epilogue thread) for all T steps;
//   * W_hh (256 KB in fp32 containers) does not fit next to the h tiles, so it streams from L2 every step through a 3-stage
//     TMA ring of [4 gates x 64 units] x 32-float slabs - the same boxes the per-timestep kernel uses; the ring runs ahead
//     of the recurrence (the weights depend on nothing);
//   * the input projection (+ biases) is read from the gate buffer with 16-byte loads issued before the wait for the tensor
//     core; the grad-carrying variant overwrites it with the activated gates and stores c, as the per-timestep kernel does.
// Per step the tensor pipe works 2 x 16 MMAs (M128 x N256 x K8), the XU pipe 7 MUFU per cell; the two are about balanced
// at H = 128, and only the first half's MMAs are exposed.
// ---------------------------------------------------------------------------------------------
constexpr int PH = 128;                   // hidden units
constexpr int P_HBUF = 128 * PH * 4;      // 64 KB: one h tile = 4 k-slabs of [128 rows][128 bytes]
constexpr int P_SLAB = 128 * BK * 4;      // 16 KB
constexpr int P_WST = 4 * SUN * BK * 4;   // 32 KB: one weight stage (4 gates x 64 units, one k-slab)
constexpr int P_NST = 3;
constexpr int P_KT = PH / BK;             // 4 k-slabs
constexpr int P_THREADS = 384;            // warpgroup 0: warp 0 TMA producer, warp 1 MMA issuer; warpgroups 1, 2: epilogue

struct PersistParams {
  CUtensorMap b[2];   // per direction: W_hh [4H][H], boxes of 64 rows
  CUtensorMap hs[2];  // per direction: h of that direction inside hseq as (k, gesture, t) - the TMA store of each h_t tile
  float* gates;       // CHUNK: [2][T][tiles][4H / 4][128][4]; else [2][T][B][4H]
  float* cseq;        // STORE: [2][T][B][H]
  int T, B;
  int l2pf;           // > 0: L2 prefetch distance (steps) for the gate blocks
  int* gerr;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(src)
               : "memory");
}

// STORE = 1: grad-carrying pass (the gate buffer is overwritten with the activated gates, c is stored).  CHUNK = 1: the gate
// buffer (and c) in the chunked order [t][tile][columns / 4][128 rows][4 floats] the input-projection GEMM wrote for this kernel
// - every epilogue access is a coalesced 16-byte access; BPTT (gemm_tc_lstm_bwd_kernel<1>) reads a k-slab of it as one
// contiguous 16 KB block in UMMA core-matrix order.  CHUNK = 0: row-major [t][b][columns] (kept for A/B measurements).
template <int STORE, int CHUNK>
__global__ void __launch_bounds__(P_THREADS, 1) lstm128_tc_fwd_kernel(const __grid_constant__ PersistParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_h = smem;                      // two h tiles
  uint8_t* s_w = smem + 2 * P_HBUF;         // weight ring
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_w + P_NST * P_WST);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 16);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const int m0 = blockIdx.x * SBM;
  const int T = p.T;
  const uint32_t bar0 = smem_u32(s_bar);
  auto WFULL = [&](int s) { return bar0 + 8u * s; };
  auto WEMPTY = [&](int s) { return bar0 + 8u * (P_NST + s); };
  auto ACC_FULL = [&](int h) { return bar0 + 8u * (2 * P_NST + h); };
  auto H_FULL = [&](int h) { return bar0 + 8u * (2 * P_NST + 2 + h); };  // h_t units of half h are in the operand tile
  if (tid == 0) {
    for (int s = 0; s < P_NST; ++s) { mbar_init(WFULL(s), 1); mbar_init(WEMPTY(s), 1); }
    mbar_init(ACC_FULL(0), 1);
    mbar_init(ACC_FULL(1), 1);
    mbar_init(H_FULL(0), 8);  // the 8 epilogue warps, once per half and step
    mbar_init(H_FULL(1), 8);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // Register budget: 384 threads start with 168 registers each; the first warpgroup (TMA producer, MMA issuer, two idle warps)
  // hands 112 per thread back, the two epilogue warpgroups (64 cell states + two prefetched items + the accumulator chunk per
  // thread) take 224 - at 168 the epilogue spilled ~0.5 KB per thread into an L1 that the 224 KB of shared memory leaves at
  // 32 KB, and the local-memory round trips to L2 were the largest stall of the kernel (ncu: long scoreboard).
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    if (lane == 0) {
      // weight ring: (step, half, k-slab) in the order the issuer consumes them; no step-0 product (h_{-1} = 0)
      int n = 0;
      bool ok = true;
      // the tile's input projection of one timestep is one contiguous block (256 KB for a full tile) in either layout;
      // optionally (PersistParams::l2pf = distance in steps, off by default) it is pulled into L2 ahead of the epilogue.
      // Measured with every SM busy: distance 2 doubles the DRAM reads (148 x 256 KB x 3 blocks in flight do not survive in
      // L2 until they are used) and is slower than no prefetch.
      const int tile_rows = p.B - m0 < SBM ? p.B - m0 : SBM;
      const uint32_t blk_bytes = (uint32_t)(CHUNK ? SBM : tile_rows) * 4 * PH * 4;
      auto gate_block = [&](int step) -> const uint8_t* {
        const int t = dir ? T - 1 - step : step;
        const int64_t tiles = (p.B + SBM - 1) / SBM;
        const int64_t row = CHUNK ? (((int64_t)dir * T + t) * tiles + blockIdx.x) * SBM : ((int64_t)dir * T + t) * p.B + m0;
        return reinterpret_cast<const uint8_t*>(p.gates) + row * (4 * PH * 4);
      };
      auto l2_prefetch = [&](int step) {
        if (step >= T) return;
        const uint8_t* src = gate_block(step);
        for (uint32_t off = 0; off < blk_bytes; off += 32768) {
          const uint32_t nb = blk_bytes - off < 32768 ? blk_bytes - off : 32768;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + off), "r"(nb) : "memory");
        }
      };
      if (p.l2pf > 0)
        for (int d = 1; d <= p.l2pf; ++d) l2_prefetch(d);
      for (int step = 1; step < T && ok; ++step) {
        if (p.l2pf > 0) l2_prefetch(step + p.l2pf);
        for (int half = 0; half < 2 && ok; ++half)
          for (int k = 0; k < P_KT; ++k, ++n) {
            const int s = n % P_NST;
            if (!mbar_wait(WEMPTY(s), (uint32_t)(((n / P_NST) & 1) ^ 1), s_abort, p.gerr, 81)) { ok = false; break; }
            mbar_expect_tx(WFULL(s), P_WST);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              tma_load_2d(smem_u32(s_w + s * P_WST + g * SBOX_BYTES), &p.b[dir], k * BK, g * PH + half * SUN, WFULL(s));
          }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, 256);
    int n = 0;
    bool ok = true;
    // Issue order per step: the first two k-slabs of the first accumulator half only need the first half of h_{step-1}
    // (units 0..63, published on H_FULL(0)) and run while the epilogue still works on the second half of the previous step;
    // accumulator half 1 is still being read then and is touched only after H_FULL(1).
    auto issue = [&](int half, int k0, int k1, uint32_t a0) -> bool {
      for (int k = k0; k < k1; ++k, ++n) {
        const int s = n % P_NST;
        if (!mbar_wait(WFULL(s), (uint32_t)((n / P_NST) & 1), s_abort, p.gerr, 83)) return false;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b0 = smem_u32(s_w + s * P_WST);
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks)
            mma_tf32_ss(tmem_base + (uint32_t)(half * 256), desc_sw128(a0 + k * P_SLAB + ks * 32), desc_sw128(b0 + ks * 32),
                        idesc, (k | ks) ? 1u : 0u);
          mma_commit(WEMPTY(s));
        }
        __syncwarp();
      }
      return true;
    };
    for (int step = 1; step <= T && ok; ++step) {
      const uint32_t par = (uint32_t)((step - 1) & 1);
      const uint32_t a0 = smem_u32(s_h + ((step - 1) & 1) * P_HBUF);
      if (!mbar_wait(H_FULL(0), par, s_abort, p.gerr, 82)) break;
      tc_fence_after();
      if (step < T && !issue(0, 0, 2, a0)) break;
      // h_{step-1} complete in tile (step-1) & 1 (and with it: every accumulator read of the previous step)
      if (!mbar_wait(H_FULL(1), par, s_abort, p.gerr, 85)) break;
      tc_fence_after();
      if (lane == 0) {
        // the finished tile goes to hseq (the next layer's input) by TMA: four [128 gestures x 32 units] boxes; rows beyond
        // the batch are clipped by the map.  The tile is overwritten by the epilogue of step + 1, which cannot start before
        // this warp has committed that step's first accumulator half - after the wait below.
        const int tprev = dir ? T - step : step - 1;
#pragma unroll
        for (int k = 0; k < P_KT; ++k) tma_store_3d(&p.hs[dir], k * BK, m0, tprev, a0 + k * P_SLAB);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store issued one step earlier has read its tile
      }
      __syncwarp();
      if (step == T) break;
      if (!issue(0, 2, P_KT, a0)) break;
      if (elect_one()) mma_commit(ACC_FULL(0));
      __syncwarp();
      if (!issue(1, 0, P_KT, a0)) break;
      if (elect_one()) mma_commit(ACC_FULL(1));
      __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int quarter = warp & 3, uh = (warp - 4) >> 2;
    const int rl = quarter * 32 + lane;  // row inside the tile
    const int row = m0 + rl;
    const bool rok = row < p.B;
    const int64_t r = rok ? row : 0;
    const int64_t TB = (int64_t)T * p.B;
    const int tiles = (p.B + SBM - 1) / SBM;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // gate-buffer row of this thread at sequence position t
    auto gate_ptr = [&](int t) -> float* {
      if (CHUNK) return p.gates + ((((int64_t)dir * T + t) * tiles + blockIdx.x) * (4 * PH / 4) * 128 + rl) * 4;
      return p.gates + (((int64_t)dir * T + t) * p.B + r) * 4 * PH;
    };
    // 8 items per step: item q = (half q >> 2, chunk q & 3) = units [u, u + 8) of this warp's share of the half; an item's
    // input projection (eight 16-byte loads) is requested two items ahead, across the half and the step boundary (issuing
    // the loads only after the half's proxy fence - which waits for outstanding loads - was measured slower)
    float4 pre[2][4][2];
    auto load_item = [&](const float* gp, int q, int slot) {
      const int u = (q >> 2) * SUN + uh * 32 + (q & 3) * 8;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (CHUNK) {
          const float* a = gp + (int64_t)((g * PH + u) >> 2) * 512;
          pre[slot][g][0] = ld4(a);
          pre[slot][g][1] = ld4(a + 512);
        } else {
          pre[slot][g][0] = rok ? ld4(gp + g * PH + u) : z4;
          pre[slot][g][1] = rok ? ld4(gp + g * PH + u + 4) : z4;
        }
      }
    };
    float creg[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int i = 0; i < 8; ++i) creg[a][i] = 0.f;
    float* gp = gate_ptr(dir ? T - 1 : 0);
    load_item(gp, 0, 0);
    load_item(gp, 1, 1);
    bool live = true;
#pragma unroll 1
    for (int step = 0; step < T && live; ++step) {
      const int t = dir ? T - 1 - step : step;
      float* gp_next = step + 1 < T ? gate_ptr(dir ? t - 1 : t + 1) : gp;
      // c of this thread's row at t: row-major [2][T][B][H], or chunked [2][T][tiles][H / 4][128][4] with the gate buffer
      float* cout = !STORE ? nullptr
                    : CHUNK ? p.cseq + ((((int64_t)dir * T + t) * tiles + blockIdx.x) * (PH / 4) * 128 + rl) * 4
                            : p.cseq + (int64_t)dir * TB * PH + ((int64_t)t * p.B + r) * PH;
      uint8_t* hb = s_h + (step & 1) * P_HBUF + rl * 128;  // this row inside each k-slab of the tile being produced
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int half = q >> 2, slot = q & 1;
        const int u = half * SUN + uh * 32 + (q & 3) * 8;
        if ((q & 3) == 0 && step > 0) {
          live = mbar_wait(ACC_FULL(half), (uint32_t)((step - 1) & 1), s_abort, p.gerr, 84);
          tc_fence_after();
        }
        if (!live) break;
        float a[4][8];
        if (step > 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t rr[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7])
                         : "r"(taddr + (uint32_t)(half * 256 + g * SUN + uh * 32 + (q & 3) * 8)));
#pragma unroll
            for (int i = 0; i < 8; ++i) a[g][i] = __uint_as_float(rr[i]);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[g][i] = 0.f;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 lo = pre[slot][g][0], hi = pre[slot][g][1];
          a[g][0] += lo.x; a[g][1] += lo.y; a[g][2] += lo.z; a[g][3] += lo.w;
          a[g][4] += hi.x; a[g][5] += hi.y; a[g][6] += hi.z; a[g][7] += hi.w;
        }
        if (q + 2 < 8) load_item(gp, q + 2, slot);
        else if (step + 1 < T) load_item(gp_next, q + 2 - 8, slot);
        float hv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float ig, fg, gg, og, hh, cn;
          lstm_cell_fast(a[0][i], a[1][i], a[2][i], a[3][i], creg[q][i], ig, fg, gg, og, cn, hh);
          creg[q][i] = cn;
          hv[i] = rna_tf32(hh);
          a[0][i] = ig; a[1][i] = fg; a[2][i] = gg; a[3][i] = og;
        }
        // next step's A operand (and the source of the TMA store to hseq): K-major SWIZZLE_128B - the 16-byte chunk index
        // is XOR-ed with (row & 7) inside the row's 128 bytes
        {
          uint8_t* slab = hb + (u >> 5) * P_SLAB;
          const int ch = (u & 31) >> 2;  // even
          *reinterpret_cast<float4*>(slab + ((ch ^ (rl & 7)) << 4)) = make_float4(hv[0], hv[1], hv[2], hv[3]);
          *reinterpret_cast<float4*>(slab + (((ch + 1) ^ (rl & 7)) << 4)) = make_float4(hv[4], hv[5], hv[6], hv[7]);
        }
        if (STORE && rok) {
          if (CHUNK) {
            float* cq = cout + (int64_t)(u >> 2) * 512;
            st4(cq, make_float4(creg[q][0], creg[q][1], creg[q][2], creg[q][3]));
            st4(cq + 512, make_float4(creg[q][4], creg[q][5], creg[q][6], creg[q][7]));
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float* gq = gp + (int64_t)((g * PH + u) >> 2) * 512;
              st4(gq, make_float4(a[g][0], a[g][1], a[g][2], a[g][3]));
              st4(gq + 512, make_float4(a[g][4], a[g][5], a[g][6], a[g][7]));
            }
          } else {
            st4(cout + u, make_float4(creg[q][0], creg[q][1], creg[q][2], creg[q][3]));
            st4(cout + u + 4, make_float4(creg[q][4], creg[q][5], creg[q][6], creg[q][7]));
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              st4(gp + g * PH + u, make_float4(a[g][0], a[g][1], a[g][2], a[g][3]));
              st4(gp + g * PH + u + 4, make_float4(a[g][4], a[g][5], a[g][6], a[g][7]));
            }
          }
        }
        if ((q & 3) == 3) {
          // this warp's share of h_t (units of this half) is in the operand tile and its accumulator reads are done
          tc_fence_before();
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(H_FULL(half));
        }
      }
      gp = gp_next;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// W^T image: out[d][k][n] = in[d][n][k]   (n < N rows, k < K columns), 32 x 32 tiles through shared memory
__global__ void transpose_image_kernel(const float* __restrict__ in, int64_t in_bs, float* __restrict__ out, int N, int K) {
  __shared__ float tile[32][33];
  const float* src = in + (int64_t)blockIdx.z * in_bs;
  float* dst = out + (int64_t)blockIdx.z * N * K;
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    tile[i][threadIdx.x] = (n < N && k < K) ? src[(int64_t)n * K + k] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    if (k < K && n < N) dst[(int64_t)k * N + n] = rna_tf32(tile[threadIdx.x][i]);
  }
}

// out[b][c][r] = tf32_rna(in[b][r][c])  (r < R rows of stride ld_in, c < C; out rows of stride ld_out): the K-major images of
// the weight-gradient operands (K = T * B runs along the rows of da / the layer input, the transpose of what TMA +
// tcgen05.mma kind::tf32 can take at full rate).  64 x 64 tiles, 16-byte accesses on both sides.
__global__ void __launch_bounds__(256) transpose_tf32_kernel(const float* __restrict__ in, int64_t ld_in, int64_t bs_in,
                                                             float* __restrict__ out, int64_t ld_out, int64_t bs_out, int64_t R,
                                                             int C) {
  __shared__ float tile[64][65];
  const float* src = in + (int64_t)blockIdx.z * bs_in;
  float* dst = out + (int64_t)blockIdx.z * bs_out;
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 floats each along the contiguous side
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + 16 * i;
    const int c = c0 + 4 * tx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R && c < C) v = *reinterpret_cast<const float4*>(src + r * ld_in + c);  // C % 4 == 0
    tile[ty + 16 * i][4 * tx] = v.x; tile[ty + 16 * i][4 * tx + 1] = v.y;
    tile[ty + 16 * i][4 * tx + 2] = v.z; tile[ty + 16 * i][4 * tx + 3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 16 * i;
    const int64_t r = r0 + 4 * tx;
    if (c < C && r < R) {  // R % 4 == 0
      const float4 v = make_float4(rna_tf32(tile[4 * tx][ty + 16 * i]), rna_tf32(tile[4 * tx + 1][ty + 16 * i]),
                                   rna_tf32(tile[4 * tx + 2][ty + 16 * i]), rna_tf32(tile[4 * tx + 3][ty + 16 * i]));
      *reinterpret_cast<float4*>(dst + (int64_t)c * ld_out + r) = v;
    }
  }
}

}  // namespace gtc

int transpose_tf32_launch(wgg_ctx* ctx, const float* in, int64_t ld_in, int64_t bs_in, float* out, int64_t ld_out,
                          int64_t bs_out, int64_t R, int C, int nbatch, cudaStream_t st) {
  if ((R & 3) || (C & 3) || (ld_in & 3) || (ld_out & 3) || (bs_in & 3) || (bs_out & 3) ||
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15))
    return wgg_fail(ctx, WGG_EINVAL, "transpose_tf32: operands must be 16-byte aligned with dimensions %% 4 == 0%s");
  dim3 grid((unsigned)cdiv64(R, 64), (unsigned)cdiv64(C, 64), (unsigned)nbatch);
  ProfScope prof(ctx, "transpose_tf32_kernel", st, 0.0, 8.0 * (double)R * C * nbatch, "transpose_tf32_kernel");
  gtc::transpose_tf32_kernel<<<grid, 256, 0, st>>>(in, ld_in, bs_in, out, ld_out, bs_out, R, C);
  WGG_CHECK_LAUNCH(ctx, "transpose_tf32_kernel");
  return WGG_OK;
}

// W^T images for the input-gradient GEMM: out[d][k][n] = in[d][n][k] (n < N rows, k < K columns of a row-major weight)
int transpose_image_launch(wgg_ctx* ctx, const float* in, int64_t in_bs, float* out, int N, int K, int nbatch, cudaStream_t st) {
  dim3 g((unsigned)cdiv64(K, 32), (unsigned)cdiv64(N, 32), (unsigned)nbatch);
  gtc::transpose_image_kernel<<<g, dim3(32, 8), 0, st>>>(in, in_bs, out, N, K);
  WGG_CHECK_LAUNCH(ctx, "transpose_image_kernel");
  return WGG_OK;
}

// The weight-gradient / input-gradient GEMMs of the step-by-step (scaled) LSTM path can take the tcgen05 engine when the
// transposed images are 16-byte addressable: T * B and B multiples of 4 (time shifts are column offsets of B floats).
bool lstm_wgrad_tc_usable(const wgg_ctx* ctx, int H, int64_t B, int T) {
  return ctx->math_mode >= 1 && H >= 32 && (H & 7) == 0 && (B & 3) == 0 && (int64_t)T * B >= 1024 && gtc::encode_fn() != nullptr;
}

// Can this contraction go through the tcgen05 GEMM?  (fully contiguous-K operands, 16-byte aligned rows, unit column
// stride of C, no activation, no conv window, no split-K)
bool gemm_tc_usable(const wgg_ctx* ctx, const GemmP& p) {
  if (ctx->math_mode < 1 || p.force_fp32 || p.conv_mode || p.act != ACT_NONE || p.rowsum) return false;
  // split-K (the weight gradients: K = T * B) goes through the dense partial buffer and the fixed-order reduction
  if (p.splitk > 1 && (!p.partial || p.bias || p.bias2 || p.scm != p.N)) return false;
  if (p.sak != 1 || p.sbk != 1 || p.scn != 1 || p.nbatch > 2) return false;
  if ((p.sam & 3) || (p.sbn & 3) || (p.scm & 3) || (p.N & 3) || (p.K & 3)) return false;
  if (p.M < 128 || p.N < 128 || p.K < 32) return false;  // small problems stay on the mma.sync engine
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al(p.A) || !al(p.B) || !al(p.C) || (p.bias && !al(p.bias)) || (p.bias2 && !al(p.bias2))) return false;
  if (p.splitk > 1 && !al(p.partial)) return false;
  if (p.out_chunk && (p.splitk > 1 || p.accumulate || p.chunk_B <= 0)) return false;
  if (p.nbatch == 2 && ((p.bsA & 3) || (p.bsB & 3) || (p.bsC & 3) || (p.bsBias & 3))) return false;
  return gtc::encode_fn() != nullptr;
}

int gemm_tc_launch(wgg_ctx* ctx, const GemmP& p, cudaStream_t st) {
  gtc::Params prm;
  memset(&prm, 0, sizeof(prm));
  // split-K: the caller's split count was chosen for the mma.sync tiling; this engine picks its own (one wave of 256 x 256
  // tiles, at least 8 k-slabs per split, within the partial buffer's capacity of 128 slices)
  int splits = 1;
  int64_t k_len = ((p.K + gtc::BK - 1) / gtc::BK) * gtc::BK;
  if (p.splitk > 1) {
    const int64_t tiles = cdiv64(p.N, gtc::BN) * cdiv64(p.M, gtc::BM) * p.nbatch;
    const int64_t kt = cdiv64(p.K, gtc::BK);
    int64_t want = cdiv64(ctx->sm_count, tiles);
    if (want > kt / 8) want = kt / 8;
    if (want > 128) want = 128;
    if (want < 1) want = 1;
    k_len = cdiv64(kt, want) * gtc::BK;
    splits = (int)cdiv64(p.K, k_len);  // no empty split
  }
  const bool split = p.splitk > 1;
  for (int z = 0; z < p.nbatch; ++z) {
    if (!gtc::make_map(&prm.a[z], p.A + z * p.bsA, p.M, p.K, p.sam) || !gtc::make_map(&prm.b[z], p.B + z * p.bsB, p.N, p.K, p.sbn))
      return wgg_fail(ctx, WGG_ECUDA, "gemm_tc: cuTensorMapEncodeTiled failed%s");
    prm.c[z] = split ? p.partial + (int64_t)z * splits * p.M * p.N : p.C + z * p.bsC;
    prm.bias[z] = p.bias ? p.bias + z * p.bsBias : nullptr;
    prm.bias2[z] = p.bias2 ? p.bias2 + z * p.bsBias : nullptr;
  }
  prm.M = (int)p.M; prm.N = (int)p.N; prm.K = (int)p.K; prm.ldc = split ? (int)p.N : (int)p.scm;
  prm.accumulate = split ? 0 : p.accumulate;
  prm.splits = splits; prm.k_len = (int)k_len;
  prm.out_chunk = p.out_chunk; prm.chunk_B = (int)p.chunk_B; prm.chunk_tiles = (int)cdiv64(p.chunk_B > 0 ? p.chunk_B : 1, 128);
  prm.gerr = ctx->async_err;
  constexpr size_t smem = (size_t)2 * gtc::NST * gtc::TILE_BYTES + 64 + 16 + gtc::BN * 4 + 1024;
  if (!wgg_smem_ok(ctx, gtc::gemm_tc_nt_kernel, smem)) return wgg_fail(ctx, WGG_ECUDA, "gemm_tc_nt_kernel: cannot reserve shared memory%s");
  dim3 grid((unsigned)cdiv64(p.N, gtc::BN), (unsigned)cdiv64(p.M, gtc::BM), (unsigned)(p.nbatch * splits));
  {
    ProfScope prof(ctx, "gemm_tc_nt_kernel", st, 2.0 * p.M * (double)p.N * p.K * p.nbatch,
                   4.0 * p.nbatch * ((double)p.M * p.K + (double)p.N * p.K + (double)p.M * p.N * (p.accumulate ? 2 : 1)),
                   p.tag ? p.tag : "gemm_tc_nt_kernel");
    gtc::gemm_tc_nt_kernel<<<grid, gtc::THREADS, smem, st>>>(prm);
    WGG_CHECK_LAUNCH(ctx, "gemm_tc_nt_kernel");
  }
  if (split)
    WGG_TRY(reduce_partials_launch(ctx, p.partial, splits, p.M * p.N, p.nbatch, (int64_t)splits * p.M * p.N, p.C, nullptr,
                                   p.bsC, p.accumulate, st));
  return WGG_OK;
}

namespace gtc {

// 2-D map over a [rows, K] matrix (row stride ld floats) with boxes of box_rows x 32 floats
bool make_map_rows(CUtensorMap* m, const float* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 3-D map (k, gesture, t) over a time-major activation: K columns of each row, row stride ld floats, B rows per timestep
bool make_map_time(CUtensorMap* m, const float* base, int64_t K, int64_t B, int64_t T, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)B, (cuuint64_t)T};
  const cuuint64_t gstride[2] = {(cuuint64_t)ld * 4, (cuuint64_t)B * ld * 4};
  const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)SBM, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Launch of one timestep kernel with programmatic stream serialisation (the kernels call griddepcontrol.wait before they
// touch anything the previous launch wrote); WGG_PDL=0 falls back to ordinary stream order.
template <class P>
int launch_step(wgg_ctx* ctx, void (*kernel)(const P), dim3 grid, size_t smem, cudaStream_t st, const P& prm, const char* name,
                int cluster_y = 1) {
  static const bool pdl = [] { const char* e = getenv("WGG_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_y > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = 1; at[n].val.clusterDim.y = (unsigned)cluster_y; at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kernel, prm);
  return wgg_check_launch(ctx, name);
}

}  // namespace gtc

// The fused per-step kernels take the scaled recurrence in the tensor-core math modes when every 16-byte access they make
// is aligned (H a multiple of 8) and the driver exposes cuTensorMapEncodeTiled.
bool lstm_step_tc_usable(const wgg_ctx* ctx, int H, const float* gates, const float* hseq, const float* lp, int64_t off_whh,
                         int64_t dir_stride) {
  if (ctx->math_mode < 1 || (H & 7) || H < 32) return false;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al(gates) || !al(hseq) || !al(lp + off_whh) || (dir_stride & 3)) return false;
  return gtc::encode_fn() != nullptr;
}

int lstm_step_tc_forward(wgg_ctx* ctx, int H, float* gates, const float* lp, int64_t dir_stride, int64_t off_whh, float* hseq,
                         float* cseq, float* cstate, int T, int64_t B, int store, int chunked, cudaStream_t st) {
  gtc::StepFwdParams prm;
  memset(&prm, 0, sizeof(prm));
  for (int d = 0; d < 2; ++d) {
    if (!gtc::make_map_time(&prm.a[d], hseq + d * H, H, B, T, 2 * H) ||
        !gtc::make_map_rows(&prm.b[d], lp + off_whh + d * dir_stride, 4 * H, H, H, gtc::SUN))
      return wgg_fail(ctx, WGG_ECUDA, "lstm_step_tc_forward: cuTensorMapEncodeTiled failed%s");
  }
  prm.gates = gates; prm.hseq = hseq; prm.cseq = cseq; prm.cstate = cstate;
  prm.T = T; prm.B = (int)B; prm.H = H; prm.store = store; prm.gerr = ctx->async_err;
  static const int nst = [] { const char* e = getenv("WGG_STEP_NST"); const int v = e ? atoi(e) : 3; return v == 2 || v == 4 ? v : 3; }();
  const size_t smem = (size_t)nst * (gtc::SA_BYTES + gtc::SB_BYTES) + 128 + 16 + 1024;
  // WGG_STEP_CLUSTER=2: pairs of row tiles share the weight slabs by TMA multicast (3-stage ring, at least two row tiles)
  static const int cl_env = [] { const char* e = getenv("WGG_STEP_CLUSTER"); return e ? atoi(e) : 1; }();
  const int cl = (cl_env == 2 && nst == 3 && cdiv64(B, gtc::SBM) >= 2) ? 2 : 1;
  void (*kernel)(const gtc::StepFwdParams) =
      cl == 2 ? (chunked ? gtc::gemm_tc_lstm_fwd_kernel<3, 1, 2> : gtc::gemm_tc_lstm_fwd_kernel<3, 0, 2>)
      : chunked ? (nst == 2 ? gtc::gemm_tc_lstm_fwd_kernel<2, 1, 1> : nst == 4 ? gtc::gemm_tc_lstm_fwd_kernel<4, 1, 1> : gtc::gemm_tc_lstm_fwd_kernel<3, 1, 1>)
                : (nst == 2 ? gtc::gemm_tc_lstm_fwd_kernel<2, 0, 1> : nst == 4 ? gtc::gemm_tc_lstm_fwd_kernel<4, 0, 1> : gtc::gemm_tc_lstm_fwd_kernel<3, 0, 1>);
  if (!wgg_smem_ok(ctx, kernel, smem))
    return wgg_fail(ctx, WGG_ECUDA, "gemm_tc_lstm_fwd_kernel: cannot reserve shared memory%s");
  // row tiles padded to whole clusters: the extra CTA fetches its share of the weights, multiplies zero-filled rows, stores nothing
  dim3 grid((unsigned)cdiv64(H, gtc::SUN), (unsigned)(cdiv64(cdiv64(B, gtc::SBM), cl) * cl), 2);
  for (int step = 0; step < T; ++step) {
    prm.step = step;
    ProfScope prof(ctx, "gemm_tc_lstm_fwd_kernel", st, step > 0 ? 2.0 * B * 4.0 * H * H * 2 : 0.0,
                   4.0 * 2 * ((double)B * (step > 0 ? H : 0) + (double)B * 4 * H * (store ? 2 : 1) + 3.0 * B * H),
                   "gemm_tc_lstm_fwd_kernel");
    WGG_TRY(gtc::launch_step(ctx, kernel, grid, smem, st, prm, "gemm_tc_lstm_fwd_kernel", cl));
  }
  return WGG_OK;
}

// gen_hidden_dim = 128 in the tensor-core modes: the persistent kernel above.  WGG_LSTM128_PERSIST = 0 keeps the
// per-timestep launches everywhere, 1 (default) uses the persistent kernel where the gate buffer is chunked (no-grad passes:
// sampling and the critic phase's generations, 10 of the 12 generator batches of a step), 2 also for the row-major
// grad-carrying passes - measured slower there than the per-timestep kernels (lane-strided 16-byte accesses to the gate
// buffer: 134 vs 108 ms per H = 128 / T = 256 / B = 1024 step), kept for A/B measurements.
static int lstm128_mode() {
  static const int m = [] { const char* e = getenv("WGG_LSTM128_PERSIST"); return e ? atoi(e) : 1; }();
  return m;
}
bool lstm128_persist_usable(const wgg_ctx* ctx, int H, const float* gates, const float* hseq, const float* lp, int64_t off_whh,
                            int64_t dir_stride) {
  return lstm128_mode() >= 1 && H == gtc::PH && lstm_step_tc_usable(ctx, H, gates, hseq, lp, off_whh, dir_stride);
}
bool lstm128_persist_rowmajor() { return lstm128_mode() >= 2; }

int lstm128_persist_forward(wgg_ctx* ctx, float* gates, const float* lp, int64_t dir_stride, int64_t off_whh, float* hseq,
                            float* cseq, int T, int64_t B, int store, int chunked, cudaStream_t st) {
  constexpr int H = gtc::PH;
  gtc::PersistParams prm;
  memset(&prm, 0, sizeof(prm));
  for (int d = 0; d < 2; ++d)
    if (!gtc::make_map_rows(&prm.b[d], lp + off_whh + d * dir_stride, 4 * H, H, H, gtc::SUN) ||
        !gtc::make_map_time(&prm.hs[d], hseq + d * H, H, B, T, 2 * H))
      return wgg_fail(ctx, WGG_ECUDA, "lstm128_persist_forward: cuTensorMapEncodeTiled failed%s");
  prm.gates = gates; prm.cseq = cseq; prm.T = T; prm.B = (int)B; prm.gerr = ctx->async_err;
  // L2 prefetch of the next step's gate block: pays when at most ~half of the SMs run this kernel (measured 7.9 vs 9.1 ms per
  // 4-layer forward at 64 CTAs), hurts with every SM busy (14.6 vs 10.6 ms at 148 CTAs: the blocks do not survive in L2)
  static const int l2pf_env = [] { const char* e = getenv("WGG_LSTM128_L2PF"); return e ? atoi(e) : -1; }();
  prm.l2pf = l2pf_env >= 0 ? l2pf_env : (2 * cdiv64(B, gtc::SBM) <= 64 ? 1 : 0);
  constexpr size_t smem = (size_t)2 * gtc::P_HBUF + gtc::P_NST * gtc::P_WST + 128 + 16 + 1024;
  void (*kernel)(const gtc::PersistParams) =
      store ? (chunked ? gtc::lstm128_tc_fwd_kernel<1, 1> : gtc::lstm128_tc_fwd_kernel<1, 0>)
            : (chunked ? gtc::lstm128_tc_fwd_kernel<0, 1> : gtc::lstm128_tc_fwd_kernel<0, 0>);
  if (!wgg_smem_ok(ctx, kernel, smem))
    return wgg_fail(ctx, WGG_ECUDA, "lstm128_tc_fwd_kernel: cannot reserve shared memory%s");
  dim3 grid((unsigned)cdiv64(B, gtc::SBM), 2);
  const double TB = (double)T * B;
  ProfScope prof(ctx, "lstm128_tc_fwd_kernel", st, 2.0 * (T - 1) * (double)B * 4.0 * H * H * 2,
                 4.0 * 2 * (TB * 4 * H * (store ? 2 : 1) + TB * H * (store ? 2 : 1)), "lstm128_tc_fwd_kernel");
  kernel<<<grid, gtc::P_THREADS, smem, st>>>(prm);
  WGG_CHECK_LAUNCH(ctx, "lstm128_tc_fwd_kernel");
  return WGG_OK;
}

// scratch: dcs [2][B padded to tiles][H] | W_hh^T images [2][H][4H]
int64_t lstm_step_tc_bwd_scratch_floats(int H, int64_t B) { return 2 * ((B + 127) / 128 * 128) * (int64_t)H + 8 * (int64_t)H * H + 64; }

int lstm_step_tc_backward(wgg_ctx* ctx, int H, float* gates, const float* cseq, const float* lp, int64_t dir_stride,
                          int64_t off_whh, const float* dh_out, float* scratch, int T, int64_t B, int chunked, cudaStream_t st) {
  float* dcs = scratch;
  float* wt = scratch + ((2 * ((B + 127) / 128 * 128) * (int64_t)H + 3) & ~(int64_t)3);  // keep the image 16-byte aligned
  if (reinterpret_cast<uintptr_t>(wt) & 15) wt += 4 - ((reinterpret_cast<uintptr_t>(wt) & 15) >> 2);
  WGG_TRY(transpose_image_launch(ctx, lp + off_whh, dir_stride, wt, 4 * H, H, 2, st));
  const int64_t TB = (int64_t)T * B;
  gtc::StepBwdParams prm;
  memset(&prm, 0, sizeof(prm));
  for (int d = 0; d < 2; ++d) {
    if ((!chunked && !gtc::make_map_time(&prm.a[d], gates + d * TB * 4 * H, 4 * H, B, T, 4 * H)) ||
        !gtc::make_map_rows(&prm.b[d], wt + (int64_t)d * 4 * H * H, H, 4 * H, 4 * H, gtc::SUN))
      return wgg_fail(ctx, WGG_ECUDA, "lstm_step_tc_backward: cuTensorMapEncodeTiled failed%s");
  }
  prm.gates = gates; prm.cseq = cseq; prm.dh_out = dh_out; prm.dcs = dcs;
  prm.T = T; prm.B = (int)B; prm.H = H; prm.gerr = ctx->async_err;
  constexpr size_t smem = (size_t)gtc::WNST * (gtc::SA_BYTES + gtc::WB_BYTES) + 128 + 16 + 1024;
  void (*kernel)(const gtc::StepBwdParams) = chunked ? gtc::gemm_tc_lstm_bwd_kernel<1> : gtc::gemm_tc_lstm_bwd_kernel<0>;
  if (!wgg_smem_ok(ctx, kernel, smem))
    return wgg_fail(ctx, WGG_ECUDA, "gemm_tc_lstm_bwd_kernel: cannot reserve shared memory%s");
  dim3 grid((unsigned)cdiv64(H, gtc::SUN), (unsigned)cdiv64(B, gtc::SBM), 2);
  for (int step = T - 1; step >= 0; --step) {
    prm.step = step;
    ProfScope prof(ctx, "gemm_tc_lstm_bwd_kernel", st, step < T - 1 ? 2.0 * B * 4.0 * H * H * 2 : 0.0,
                   4.0 * 2 * ((double)B * 4 * H * (step < T - 1 ? 3 : 2) + 5.0 * B * H), "gemm_tc_lstm_bwd_kernel");
    WGG_TRY(gtc::launch_step(ctx, kernel, grid, smem, st, prm, "gemm_tc_lstm_bwd_kernel"));
  }
  return WGG_OK;
}

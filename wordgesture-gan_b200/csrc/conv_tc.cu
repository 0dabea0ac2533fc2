// Conv1d layers of the TemporalDiscriminator on tcgen05 tensor cores: one sample = one 128-row MMA tile.
//
// Replaces the cuDNN conv1d forward / backward-data / backward-weight calls behind
// TemporalDiscriminator.forward / get_all_features (src/gan/models.py:270-277,309-311,335-339) in "tf32" mode.
// The sequence length T = 128 equals the UMMA M dimension, so a conv over one gesture is a single accumulator
// tile D[t, co]:
//     D[t, co] = sum_{tap, ci} act[t + tap - pad, ci] * W[co, ci, tap]
// Sequences of T = 128 h points (h "halves": T = 256 for the scaled regime of BASELINE configs[3]) are h such tiles; a
// tile's input rows then include the neighbouring tile's first / last rows as halo (the work item is (gesture, tile)).
// Activations of one sample are stored channel-chunked: [C/4 chunks][T rows][4 floats].  In shared memory each
// chunk carries 2 zero rows of padding in front (and spare rows behind), rows are 16 B apart, so the K-major A
// operand of tap `j` is simply the same tile with its start address advanced by j*16 bytes - the sliding window
// costs nothing and needs no boundary masking.  K per MMA is 8 = two 16-byte chunks: two channel chunks of one tap
// (LBO = chunk stride) or, for the 3-channel input layer, two consecutive taps of the single chunk (LBO = 16 B).
// The weight gradient uses the SAME tiles as MN-major operands (M = co from the d(pre-activation) tile, N = ci from
// the shifted input tile, K = time), accumulating over all samples of a persistent CTA in TMEM; bias gradients fall
// out of one extra N=8 MMA against a tile of ones.
// Numerics: TF32 operands (round-to-nearest on write), fp32 accumulation - the numerics of the reference's own
// CUDA path (cuDNN TF32 convs; measured deviation from fp64 in profiles/r01_ref_cuda_precision.json).
#include <stdlib.h>

#include "tc_common.cuh"

using namespace tcu;

namespace ctc {

constexpr int T = 128;                 // rows of one MMA tile (UMMA M); the sequence length is a.T = 128 * a.halves
constexpr int ROWS_S = 136;            // rows per chunk in shared memory: 2 pad + 128 + 6 spare (max tap shift 5)
constexpr int CS = ROWS_S * 16;        // chunk stride in shared memory (bytes)
constexpr int PAD_ROWS = 2;
constexpr int NTHREADS = 320;          // warp 0 producer, warp 1 MMA, warps 2..9 epilogue (TMEM quarter = warp % 4, column half)
constexpr int NST = 3;                 // input ring depth (two gestures' loads in flight behind the one being multiplied)

struct FwdArgs {
  const float* in;        // [B][CinC][T][4]
  const float* wimg;      // [Kchunks][N/8][8][4] (TF32)
  const float* bias;      // [N] or null
  float* out;             // mode 0/1: [B][N/4][T][4]; mode 2: (B,T,3) row-major
  const float* act_lower; // mode 1: activation of the layer below (same layout as out) for LeakyReLU backward
  const float* dfeat;     // mode 1: optional feature-matching gradient w.r.t. that activation (same layout)
  int64_t B;
  int CinC, taps, taps_p, tap_row0, N, mode;
  int Tseq, halves;       // sequence length = 128 * halves; work items are (gesture, 128-row tile)
  int* gerr;
};

// ---------------------------------------------------------------------------------------------
// forward-type conv: mode 0 = conv + bias + LeakyReLU; mode 1 = backward-data fused with LeakyReLU backward
// (produces d(pre-activation) of the layer below); mode 2 = backward-data into the (B,T,3) input gradient.
// persistent: grid = min(B, #SM); each CTA loops over samples with a 2-deep input ring.
// ---------------------------------------------------------------------------------------------
// MULTI = false: T = 128, one tile per gesture (the default model; sequence length and item decoding fold to constants).
template <bool MULTI>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc_fwd_kernel(FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int Tseq = MULTI ? a.Tseq : T, halves = MULTI ? a.halves : 1;
  const int Kchunks = a.CinC == 1 ? a.taps_p : a.taps * a.CinC;
  const int WCS = (a.N / 8) * 128;  // weight image chunk stride
  uint8_t* s_w = smem;
  uint8_t* s_in = s_w + ((Kchunks * WCS + 1023) / 1024) * 1024;
  const int in_bytes = a.CinC * CS;
  float* s_bias = reinterpret_cast<float*>(s_in + NST * in_bytes);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 64);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * NST + 2);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  auto BAR_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_EMPTY = [&](int s) { return bar0 + 8u * (NST + s); };
  const uint32_t BAR_ACC_FULL = bar0 + 8u * (2 * NST), BAR_ACC_EMPTY = bar0 + 8u * (2 * NST + 1);

  {
    const float4* src = reinterpret_cast<const float4*>(a.wimg);
    float4* dst = reinterpret_cast<float4*>(s_w);
    for (int i = tid; i < Kchunks * WCS / 16; i += NTHREADS) dst[i] = __ldg(src + i);
    float4* zin = reinterpret_cast<float4*>(s_in);
    for (int i = tid; i < NST * in_bytes / 16; i += NTHREADS) zin[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 64; i += NTHREADS) s_bias[i] = (a.bias && i < a.N) ? __ldg(a.bias + i) : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(BAR_FULL(s), 1); mbar_init(BAR_EMPTY(s), 1); }
    mbar_init(BAR_ACC_FULL, 1);
    mbar_init(BAR_ACC_EMPTY, 8);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 64);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t idesc = make_idesc(128, a.N);

  const int64_t items = a.B * halves;
  if (warp == 0) {
    if (lane == 0) {
      int n = 0;
      const int pad = PAD_ROWS - a.tap_row0;                               // rows of halo in front of a tile
      const int hi_halo = (a.CinC == 1 ? a.taps_p : a.taps) - 1 - pad;     // ... and behind it
      const int64_t chunk_g = (int64_t)Tseq * 16;                          // one channel chunk of a gesture in HBM
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
        const int st = n % NST;
        const int64_t b = MULTI ? it / halves : it;
        const int h = MULTI ? (int)(it % halves) : 0;
        if (!mbar_wait(BAR_EMPTY(st), ((n / NST) & 1) ^ 1, s_abort, a.gerr, 11)) break;
        // sequence rows [r_lo, r_hi) land at shared row tap_row0 + (r_lo - (128 h - pad)); rows outside the sequence stay zero
        const int r_lo = h == 0 ? 0 : T * h - pad;
        int r_hi = T * h + T + (hi_halo > 0 ? hi_halo : 0);
        if (r_hi > Tseq) r_hi = Tseq;
        const int row0 = a.tap_row0 + (r_lo - (T * h - pad));
        const uint32_t nbytes = (uint32_t)(r_hi - r_lo) * 16;
        if (MULTI) {
          // a stage is reused by tiles of either kind: the rows this tile leaves unwritten (before the sequence start /
          // past its end) may hold another tile's halo - clear them (generic proxy, then the async-proxy fence)
          uint8_t* sb = s_in + st * in_bytes;
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (h == 0)
            for (int q = 0; q < a.CinC; ++q)
              for (int r = a.tap_row0; r < PAD_ROWS; ++r) *reinterpret_cast<float4*>(sb + q * CS + r * 16) = z4;
          if (h == halves - 1)
            for (int q = 0; q < a.CinC; ++q)
              for (int r = PAD_ROWS + T; r < PAD_ROWS + T + hi_halo; ++r) *reinterpret_cast<float4*>(sb + q * CS + r * 16) = z4;
          fence_async_smem();
        }
        mbar_expect_tx(BAR_FULL(st), (uint32_t)a.CinC * nbytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.in) + b * (int64_t)a.CinC * chunk_g + (int64_t)r_lo * 16;
        for (int q = 0; q < a.CinC; ++q)
          bulk_g2s(smem_u32(s_in + st * in_bytes + q * CS + row0 * 16), src + (int64_t)q * chunk_g, nbytes, BAR_FULL(st));
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues (operands stay in uniform registers)
    int n = 0;
    const uint32_t wb = smem_u32(s_w);
    const uint64_t bd0 = make_desc(wb, WCS, 128);
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int st = n % NST;
      if (!mbar_wait(BAR_FULL(st), (n / NST) & 1, s_abort, a.gerr, 12)) break;
      if (!mbar_wait(BAR_ACC_EMPTY, (n & 1) ^ 1, s_abort, a.gerr, 13)) break;
      tc_fence_after();
      const uint32_t ab = smem_u32(s_in + st * in_bytes);
      if (elect_one()) {
        uint32_t first = 0;
        if (a.CinC == 1) {
          const uint64_t ad0 = make_desc(ab + a.tap_row0 * 16, 16, 128);
          for (int jp = 0; jp < a.taps_p / 2; ++jp) {
            mma_tf32_ss(tmem_base, ad0 + (uint64_t)(2 * jp), bd0 + (uint64_t)(2 * jp * (WCS >> 4)), idesc, first);
            first = 1;
          }
        } else {
          uint64_t ad_j = make_desc(ab + a.tap_row0 * 16, CS, 128), bd = bd0;
          const uint64_t a_step = (uint64_t)((2 * CS) >> 4), b_step = (uint64_t)((2 * WCS) >> 4);
          for (int j = 0; j < a.taps; ++j, ++ad_j) {
            uint64_t ad = ad_j;
#pragma unroll 8
            for (int p = 0; p < a.CinC / 2; ++p, ad += a_step, bd += b_step) {
              mma_tf32_ss(tmem_base, ad, bd, idesc, first);
              first = 1;
            }
          }
        }
        mma_commit(BAR_EMPTY(st));
        mma_commit(BAR_ACC_FULL);
      }
      __syncwarp();
    }
  } else {
    // epilogue: 8 warps; TMEM quarter = warp % 4 (rows = time), the two warps of a quarter split the output channels
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int tl = quarter * 32 + lane;                  // row inside the tile
    const bool split = a.N >= 32;
    const int ncol = split ? a.N / 2 : a.N;             // columns handled by this warp (0 work for half 1 if !split)
    const int c0 = split ? half * ncol : 0;
    const bool active = split || half == 0;
    int n = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int64_t b = MULTI ? it / halves : it;
      const int t = (MULTI ? (int)(it % halves) * T : 0) + tl;   // row inside the sequence
      if (!mbar_wait(BAR_ACC_FULL, n & 1, s_abort, a.gerr, 14)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
      float v[32];
      if (active) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (cc * 16 < ncol) {  // warp-uniform
            float r[16];
            tmem_ld16(taddr + cc * 16, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[cc * 16 + i] = r[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_ACC_EMPTY);
      if (!active) continue;
      if (a.mode == 2) {
        float* o = a.out + (b * Tseq + t) * 3;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
      } else {
        const int64_t cbase = (b * (a.N / 4) + c0 / 4) * Tseq + t;
        float4* o4 = reinterpret_cast<float4*>(a.out) + cbase;
        const float4* y4 = a.mode == 1 ? reinterpret_cast<const float4*>(a.act_lower) + cbase : nullptr;
        const float4* f4 = (a.mode == 1 && a.dfeat) ? reinterpret_cast<const float4*>(a.dfeat) + cbase : nullptr;
        if (a.mode == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q >= ncol / 4) break;
            float x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = leaky_f(v[4 * q + i] + s_bias[c0 + 4 * q + i]);
            o4[(int64_t)q * Tseq] = make_float4(rna_tf32(x[0]), rna_tf32(x[1]), rna_tf32(x[2]), rna_tf32(x[3]));
          }
        } else {
          // LeakyReLU backward (+ feature-matching gradient injection): issue all loads first, then the math
          float4 ys[8], fs[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q < ncol / 4) {
              ys[q] = __ldg(y4 + (int64_t)q * Tseq);
              fs[q] = f4 ? __ldg(f4 + (int64_t)q * Tseq) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q >= ncol / 4) break;
            float x[4] = {v[4 * q] + fs[q].x, v[4 * q + 1] + fs[q].y, v[4 * q + 2] + fs[q].z, v[4 * q + 3] + fs[q].w};
            x[0] = ys[q].x > 0.f ? x[0] : kLeak * x[0];
            x[1] = ys[q].y > 0.f ? x[1] : kLeak * x[1];
            x[2] = ys[q].z > 0.f ? x[2] : kLeak * x[2];
            x[3] = ys[q].w > 0.f ? x[3] : kLeak * x[3];
            o4[(int64_t)q * Tseq] = make_float4(rna_tf32(x[0]), rna_tf32(x[1]), rna_tf32(x[2]), rna_tf32(x[3]));
          }
        }
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
// weight gradient: G[co][(tap, ci)] = sum_{b, t} dpre[b][t][co] * in[b][t + tap - pad][ci];  db[co] = sum dpre.
// Both operands are MN-major (M = co, N = ci, K = time).  For TF32 the only MN-major shared-memory layout the
// tensor core accepts is SWIZZLE_128B_BASE32B (verified on hardware, scripts/probe/umma_probe2.cu): a tile of
// [32-channel blocks][time rows][128 B] whose four 32-byte pieces of a row are XOR-ed with (row & 3) on ABSOLUTE
// address bits - so a start address advanced by whole rows still addresses a valid operand, which is how the tap
// shift of the sliding window is expressed.  Producer warps fill the tiles from the channel-chunked HBM layout
// with 16-byte cp.async copies (swizzle applied by the writer); the 3-channel input layer gets an explicit
// [t][tap][4] im2col tile instead.  Accumulators stay in TMEM over all samples of the persistent CTA (M = 64:
// rows sit in lanes 0..15 of each 32-lane quarter); the bias gradient is one extra N = 8 MMA against ones.
// ---------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* dpre;  // [B][CoutC][T][4]
  const float* in;    // [B][CinC][T][4]
  float* partial;     // [grid][64][ncols]
  int64_t B;
  int CoutC, CinC, taps, pad, ncols;
  int Tseq, halves;   // sequence length = 128 * halves; work items are (gesture, 128-row tile) - K = the tile's time steps
  int* gerr;
};

constexpr int W_ROWS = 136;                 // 2 zero rows + 128 + spare
constexpr int W_BLK = W_ROWS * 128;         // bytes per 32-channel block (17408 = 17 * 1024)
constexpr int W_THREADS = 160;              // warp 0: MMA issuer; warps 1..4: tile producers, then TMEM read-out

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  // MN-major SWIZZLE_128B_BASE32B: LBO = stride between 32-element MN blocks, SBO = 512 B between 4-row K atoms
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((W_BLK >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// byte offset inside a W tile of the 16-byte group (row r, channel chunk q)   [chunk = 4 channels]
__device__ __forceinline__ uint32_t w_off(int r, int q) {
  return (uint32_t)((q >> 3) * W_BLK + r * 128 + ((((q & 7) >> 1) ^ (r & 3)) << 5) + ((q & 1) << 4));
}

// First conv layer (Cin <= 4): im2col rows [x4[t-pad], ..., x4[t-pad+taps-1], 0...] make one 32-wide B block, so a
// sample is 16 k-steps of one M = 64 (co) x N = 32 MMA plus the N = 8 bias MMA.  The kernel is bound by load latency,
// not by the tensor pipe: four stages, each owned by one producer warp, keep four samples in flight.
constexpr int W1_NST = 4;
constexpr int W1_STAGE = 3 * W_BLK;  // dpre (2 blocks) + im2col rows (1 block)

template <bool MULTI>
__global__ void __launch_bounds__(W_THREADS, 1) conv_tc_wgrad_kernel(WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int Tseq = MULTI ? a.Tseq : T, halves = MULTI ? a.halves : 1;
  uint8_t* s_one = smem + W1_NST * W1_STAGE;  // 8 rows of ones
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_one + 1024);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * W1_NST + 1);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  auto BAR_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_EMPTY = [&](int s) { return bar0 + 8u * (W1_NST + s); };
  const uint32_t BAR_DONE = bar0 + 8u * (2 * W1_NST);
  {
    float4* z = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < W1_NST * W1_STAGE / 16; i += W_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* o = reinterpret_cast<float4*>(s_one);
    for (int i = tid; i < 1024 / 16; i += W_THREADS) o[i] = make_float4(1.f, 1.f, 1.f, 1.f);
  }
  if (tid == 0) {
    for (int s = 0; s < W1_NST; ++s) { mbar_init(BAR_FULL(s), 32); mbar_init(BAR_EMPTY(s), 1); }
    mbar_init(BAR_DONE, 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 64);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  constexpr int Ktot = 32;

  if (warp == 0) {
    // MMA issuer: warp-uniform loop, one elected lane issues
    int n = 0;
    bool ok = true;
    const uint64_t od = make_desc_mn(smem_u32(s_one));
    const uint32_t id_tap = make_idesc(64, 32, 1, 1);
    const uint32_t id_one = make_idesc(64, 8, 1, 1);
    const int64_t items = a.B * halves;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int st = n % W1_NST;
      if (!mbar_wait(BAR_FULL(st), (n / W1_NST) & 1, s_abort, a.gerr, 22)) { ok = false; break; }
      tc_fence_after();
      const uint32_t dp = smem_u32(smem + st * W1_STAGE);
      const uint64_t ad0 = make_desc_mn(dp + PAD_ROWS * 128), bd0 = make_desc_mn(dp + 2 * W_BLK + PAD_ROWS * 128);
      if (elect_one()) {
#pragma unroll 4
        for (int ks = 0; ks < T / 8; ++ks) {
          const uint32_t acc = (n | ks) ? 1u : 0u;
          mma_tf32_ss(tmem_base, ad0 + (uint64_t)(ks * 64), bd0 + (uint64_t)(ks * 64), id_tap, acc);
          mma_tf32_ss(tmem_base + (uint32_t)Ktot, ad0 + (uint64_t)(ks * 64), od, id_one, acc);
        }
        mma_commit(BAR_EMPTY(st));
      }
      __syncwarp();
    }
    if (ok && elect_one()) mma_commit(BAR_DONE);
  } else {
    // ---- producer warp g fills stage g for samples g, g + 4, ...: swizzled 16-byte cp.async copies ----
    const int g = warp - 1;
    int k = 0;
    bool ok = true;
    const uint32_t dp = smem_u32(smem + g * W1_STAGE);
    const uint32_t in = dp + 2 * W_BLK;
    const int64_t items = a.B * halves;
    for (int64_t it = blockIdx.x + (int64_t)g * gridDim.x; it < items; it += (int64_t)W1_NST * gridDim.x, ++k) {
      const int64_t b = MULTI ? it / halves : it;
      const int t_base = MULTI ? (int)(it % halves) * T : 0;   // first sequence row of this tile
      if (!mbar_wait(BAR_EMPTY(g), (k & 1) ^ 1, s_abort, a.gerr, 21)) { ok = false; break; }
      const float4* sd = reinterpret_cast<const float4*>(a.dpre) + b * (int64_t)a.CoutC * Tseq + t_base;
      // lane -> (4 consecutive rows) x (8 consecutive chunks): 64-byte global segments, conflict-free 512-byte smem rows
      for (int i = lane; i < a.CoutC * T; i += 32) {
        const int q = ((i >> 3) / T) * 8 + (i & 7), t = (i >> 3) % T;
        if (q < a.CoutC) cp_async16(dp + w_off(t + PAD_ROWS, q), sd + (int64_t)q * Tseq + t);
      }
      const float4* si = reinterpret_cast<const float4*>(a.in) + b * (int64_t)Tseq;
      // row t of the im2col tile = [x4[t-pad], x4[t-pad+1], ..., x4[t-pad+taps-1], 0...]: chunk index = tap
      for (int i = lane; i < 8 * T; i += 32) {
        const int j = i & 7, t = i >> 3;
        const int ts = t_base + t + j - a.pad;
        if (j < a.taps && ts >= 0 && ts < Tseq) cp_async16(in + w_off(t + PAD_ROWS, j), si + ts);
        else if (MULTI && j < a.taps)  // a stage serves tiles of either kind: what this one skips must read as zero
          asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(in + w_off(t + PAD_ROWS, j)), "f"(0.f) : "memory");
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      fence_async_smem();
      mbar_arrive(BAR_FULL(g));
    }
    // ---- read-out: M = 64 accumulator rows live in lanes 0..15 of each TMEM quarter ----
    if (ok && mbar_wait(BAR_DONE, 0, s_abort, a.gerr, 23)) {
      tc_fence_after();
      const int quarter = warp & 3;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      float* dst = a.partial + ((int64_t)blockIdx.x * 64 + quarter * 16 + lane) * a.ncols;
      for (int c0 = 0; c0 < a.ncols; c0 += 8) {
        float r[8];
        tmem_ld8(taddr + c0, r);
        if (lane < 16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[c0 + i] = r[i];
        }
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
// Weight gradient, 64-input-channel layers, K-MAJOR formulation (an earlier MN-major version of this kernel - operands
// read in HBM order - was bound by the MN-major TF32 operand fetch, roughly a quarter of the K-major rate: ~180 clk per
// M128 x N64 x K8 MMA).  Gt[(tap, ci)][co] = sum_t x[t + tap - pad][ci] * dpre[t][co] with K = time: a K-major operand row must hold
// four consecutive time steps in 16 bytes - the transpose of the HBM layout - and a tap is a shift along K that is not
// a multiple of the 16-byte chunk.  So the operand tile holds four alignment variants X_a[ci][k] = x[k + a - pad][ci]
// (a = 0..3; tap s' uses variant s' % 4 at chunk offset s' / 4), stacked as rows: [X_0; X_1] and [X_2; X_3] are M = 128
// operands that each produce two taps per MMA.  Per gesture: one loader warp bulk-copies the two HBM tiles (2 KB per
// channel chunk, chunk stride padded to 2064 B so that the gathers below are bank-conflict-free), sixteen transposer
// warps build the K-major tiles quarter by quarter (each 16-byte group = four conflict-free 4-byte shared loads +
// one 16-byte store; the bias gradient is summed on the way), one warp issues the MMAs; accumulators stay in TMEM.
// ---------------------------------------------------------------------------------------------
constexpr int W3_THREADS = 576;        // warp 0: MMA issuer; warp 1: loader; warps 2..17: transposers, then read-out
constexpr int W3_RCS = 2064;           // raw chunk stride (2048 + 16)
constexpr int W3_RAW_T = 16 * W3_RCS;  // raw dpre tile (16 chunks)
// multi-tile sequences: the raw x tile carries W3_XPAD halo rows in front of / behind the tile's 128 rows
constexpr int W3_XPAD = 4;
constexpr int W3_RCS_XM = (T + 2 * W3_XPAD) * 16 + 16;  // 2192: stride / 4 = 4 mod 32, as 2064 (conflict-free gathers)
template <bool MULTI> struct W3Geo {
  static constexpr int XPAD = MULTI ? W3_XPAD : 0;
  static constexpr int RCS_X = MULTI ? W3_RCS_XM : W3_RCS;
  static constexpr int RAW_TX = 16 * RCS_X;             // raw x tile (16 chunks)
  static constexpr int RAW_SLOT = RAW_TX + W3_RAW_T;
};
constexpr int W3_LBO_A = 256 * 16, W3_LBO_B = 64 * 16;
constexpr int W3_KCH = 8;              // k-chunks (32 time steps) per operand stage, + 1 halo chunk for chunk offset 1
constexpr int W3_STAGE_A = (W3_KCH + 1) * W3_LBO_A, W3_STAGE = W3_STAGE_A + W3_KCH * W3_LBO_B;  // 45056 B

template <bool MULTI>
__global__ void __launch_bounds__(W3_THREADS, 1) conv_tc_wgrad3_kernel(WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int W3_XP = W3Geo<MULTI>::XPAD, W3_RCS_X = W3Geo<MULTI>::RCS_X, W3_RAW_TX = W3Geo<MULTI>::RAW_TX,
                W3_RAW_SLOT = W3Geo<MULTI>::RAW_SLOT;
  const int Tseq = MULTI ? a.Tseq : T, halves = MULTI ? a.halves : 1;
  uint8_t* s_raw = smem + 2 * W3_STAGE;
  float* s_bias = reinterpret_cast<float*>(s_raw + 2 * W3_RAW_SLOT);  // [16 warps][64]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 16 * 64);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 10);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  auto BAR_FULL = [&](int s) { return bar0 + 8u * s; };             // operand stage built (16 transposer warps)
  auto BAR_EMPTY = [&](int s) { return bar0 + 16u + 8u * s; };      // operand stage consumed (MMA commit)
  auto BAR_RAW_FULL = [&](int s) { return bar0 + 32u + 8u * s; };   // raw tiles landed
  auto BAR_RAW_EMPTY = [&](int s) { return bar0 + 48u + 8u * s; };  // raw tiles read by all transposer warps
  const uint32_t BAR_DONE = bar0 + 64u;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR_FULL(s), 16); mbar_init(BAR_EMPTY(s), 1);
      mbar_init(BAR_RAW_FULL(s), 1); mbar_init(BAR_RAW_EMPTY(s), 16);
    }
    mbar_init(BAR_DONE, 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int Cout = a.CoutC * 4;
  const int npairs = (a.taps + 1) / 2;
  const bool halo = a.taps > 4;  // taps 4, 5 read their variant one chunk further

  if (warp == 0) {
    // MMA issuer: warp-uniform loop, one elected lane issues
    const uint32_t idesc = make_idesc(128, Cout);
    int n = 0;
    bool ok = true;
    const int64_t items = a.B * halves;
    for (int64_t it = blockIdx.x; it < items && ok; it += gridDim.x) {
      for (int qd = 0; qd < T / 32; ++qd, ++n) {
        const int st = n & 1;
        if (!mbar_wait(BAR_FULL(st), (uint32_t)((n >> 1) & 1), s_abort, a.gerr, 61)) { ok = false; break; }
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem) + st * W3_STAGE;
        const uint64_t ad0 = make_desc(a0, W3_LBO_A, 128), bd0 = make_desc(a0 + W3_STAGE_A, W3_LBO_B, 128);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < W3_KCH / 2; ++ks) {
            const uint32_t acc = (n | ks) ? 1u : 0u;
            const uint64_t bd = bd0 + (uint64_t)(ks * ((2 * W3_LBO_B) >> 4));
            for (int p = 0; p < npairs; ++p) {
              const int av = (2 * p) & 3, o = (2 * p) >> 2;  // variant pair [X_av; X_av+1] at chunk offset o
              mma_tf32_ss(tmem_base + (uint32_t)(p * Cout),
                          ad0 + (uint64_t)((av * 1024 + (2 * ks + o) * W3_LBO_A) >> 4), bd, idesc, acc);
            }
          }
          mma_commit(BAR_EMPTY(st));
        }
        __syncwarp();
      }
    }
    if (ok && elect_one()) mma_commit(BAR_DONE);
  } else if (warp == 1) {
    // loader: one bulk copy per channel chunk (lanes 0..15: x rows of the tile + halo, lanes 16..31: the tile's dpre rows)
    int n = 0;
    const int64_t items = a.B * halves;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int slot = n & 1;
      const int64_t b = MULTI ? it / halves : it;
      const int t_base = MULTI ? (int)(it % halves) * T : 0;
      // x rows [r_lo, r_hi) of the sequence -> raw rows r - t_base + W3_XP (rows outside the sequence are never read:
      // the transposers test the sequence index)
      const int r_lo = t_base - W3_XP < 0 ? 0 : t_base - W3_XP;
      const int r_hi = t_base + T + W3_XP > Tseq ? Tseq : t_base + T + W3_XP;
      const uint32_t xbytes = (uint32_t)(r_hi - r_lo) * 16;
      if (!mbar_wait(BAR_RAW_EMPTY(slot), (uint32_t)(((n >> 1) & 1) ^ 1), s_abort, a.gerr, 62)) break;
      if (lane == 0) mbar_expect_tx(BAR_RAW_FULL(slot), 16u * xbytes + (uint32_t)a.CoutC * (T * 16));
      __syncwarp();
      const uint32_t dst = smem_u32(s_raw) + slot * W3_RAW_SLOT;
      if (lane < 16) {
        bulk_g2s(dst + lane * W3_RCS_X + (uint32_t)(r_lo - t_base + W3_XP) * 16,
                 reinterpret_cast<const uint8_t*>(a.in) + ((b * 16 + lane) * (int64_t)Tseq + r_lo) * 16, xbytes,
                 BAR_RAW_FULL(slot));
      } else if (lane - 16 < a.CoutC) {
        bulk_g2s(dst + W3_RAW_TX + (lane - 16) * W3_RCS,
                 reinterpret_cast<const uint8_t*>(a.dpre) + ((b * a.CoutC + (lane - 16)) * (int64_t)Tseq + t_base) * 16, T * 16,
                 BAR_RAW_FULL(slot));
      }
    }
  } else {
    const int tw = warp - 2;                 // 0..15
    const int comp = lane & 3, cl = lane >> 2;
    // static work split (no per-item decoding): warp tw builds variant av = tw / 4 for channel group cg = (tw / 2) % 2,
    // k-chunks [j0, j1) of every stage, plus ONE dpre chunk (channel group tw / 8, k-chunk tw % 8)
    const int av = tw >> 2, cg = (tw >> 1) & 1;
    const int j0 = (tw & 1) ? 4 : 0;
    const int j1 = (tw & 1) ? ((halo && av < 2) ? W3_KCH + 1 : W3_KCH) : 4;  // halo chunk only feeds [X_0; X_1] at offset 1
    const int ca = cg * 8 + cl;                         // channel chunk of the A rows; channel = 4 ca + comp
    const int ra = av * 64 + 4 * ca + comp;
    const uint32_t a_src = ca * W3_RCS_X + W3_XP * 16 + comp * 4, a_dst = (ra >> 3) * 128 + (ra & 7) * 16;
    const int bcg = tw >> 3, bj = tw & 7;
    const bool b_on = bcg * 8 < a.CoutC;
    const int cb = bcg * 8 + cl, rb = 4 * cb + comp;
    const uint32_t b_src = cb * W3_RCS + comp * 4 + 4 * bj * 16, b_dst = bj * W3_LBO_B + (rb >> 3) * 128 + (rb & 7) * 16;
    float bsum = 0.f;                        // bias partial of channel bcg*32 + lane (this warp's share of the time steps)
    int n = 0, ns = 0;
    bool ok = true;
    const int64_t items = a.B * halves;
    for (int64_t it = blockIdx.x; it < items && ok; it += gridDim.x, ++ns) {
      const int slot = ns & 1;
      const int t_base = MULTI ? (int)(it % halves) * T : 0;
      if (!mbar_wait(BAR_RAW_FULL(slot), (uint32_t)((ns >> 1) & 1), s_abort, a.gerr, 63)) { ok = false; break; }
      const uint8_t* rx = s_raw + slot * W3_RAW_SLOT;
      const uint8_t* rd = rx + W3_RAW_TX;
      for (int qd = 0; qd < T / 32; ++qd, ++n) {
        const int st = n & 1;
        if (!mbar_wait(BAR_EMPTY(st), (uint32_t)(((n >> 1) & 1) ^ 1), s_abort, a.gerr, 64)) { ok = false; break; }
        uint8_t* sa = smem + st * W3_STAGE;
        uint8_t* sb = sa + W3_STAGE_A;
        {
          int t0 = 4 * (W3_KCH * qd + j0) + av - a.pad;
          const uint8_t* src = rx + a_src;
          uint8_t* dst = sa + a_dst + j0 * W3_LBO_A;
#pragma unroll
          for (int jj = 0; jj < 5; ++jj) {
            if (j0 + jj < j1) {  // warp-uniform
              float v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int t = t0 + i;  // row relative to the tile (-pad ... 128 + halo); valid if inside the sequence
                v[i] = (unsigned)(t_base + t) < (unsigned)Tseq ? *reinterpret_cast<const float*>(src + t * 16) : 0.f;
              }
              *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
              t0 += 4;
              dst += W3_LBO_A;
            }
          }
        }
        if (b_on) {
          const uint8_t* src = rd + b_src + qd * (W3_KCH * 4 * 16);
          const float v0 = *reinterpret_cast<const float*>(src), v1 = *reinterpret_cast<const float*>(src + 16);
          const float v2 = *reinterpret_cast<const float*>(src + 32), v3 = *reinterpret_cast<const float*>(src + 48);
          bsum += (v0 + v1) + (v2 + v3);
          *reinterpret_cast<float4*>(sb + b_dst) = make_float4(v0, v1, v2, v3);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR_FULL(st));
      }
      if (!ok) break;
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_RAW_EMPTY(slot));
    }
    // bias: warp tw summed channels (tw/8)*32 + lane over k-chunks = tw%8 (mod 8): combine the eight partials of each
    // channel group in fixed order
    s_bias[tw * 32 + lane] = bsum;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    float* prow = a.partial + (int64_t)blockIdx.x * 64 * a.ncols;
    if (ok && tw < 2 && tw * 32 + lane < Cout) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += s_bias[(tw * 8 + w) * 32 + lane];
      prow[(int64_t)(tw * 32 + lane) * a.ncols + a.taps * 64] = sum;
    }
    if (ok && mbar_wait(BAR_DONE, 0, s_abort, a.gerr, 65)) {
      tc_fence_after();
      // D_p[m = tapbit*64 + ci][co] (M = 128: TMEM lane = m).  Four warps per TMEM quarter split the co columns.
      const int quarter = warp & 3, cpart = tw >> 2;
      const int m = quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const int cw = Cout / 4;  // columns per warp (16 or 8)
      for (int p = 0; p < npairs; ++p) {
        const int tap = 2 * p + (m >> 6), ci = m & 63;
        for (int c0 = cpart * cw; c0 < cpart * cw + cw; c0 += 8) {
          float r[8];
          tmem_ld8(taddr + p * Cout + c0, r);
          if (tap < a.taps) {
#pragma unroll
            for (int i = 0; i < 8; ++i) prow[(int64_t)(c0 + i) * a.ncols + tap * 64 + ci] = r[i];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// G[co][(tap,ci)] = sum_cta partial[cta][co][tap*Cin4 + ci];  db[co] += sum_cta partial[cta][co][Ktot]
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ partial, int nparts, int ncols,
                                                             int Cout, int Cin, int Cin4, int taps, int Ktot,
                                                             float* __restrict__ G, float* __restrict__ db) {
  // block = 64 outputs x 4 partial groups (group g sums partials g, g+4, ...; groups are combined in fixed order)
  __shared__ float s_part[4][64];
  const int n = Cout * taps * Cin + Cout;
  const int lane_o = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int idx = blockIdx.x * 64 + lane_o;
  int co = 0, col = 0;
  const bool valid = idx < n;
  const bool isb = idx >= Cout * taps * Cin;
  if (valid) {
    if (isb) { co = idx - Cout * taps * Cin; col = Ktot; }
    else { co = idx / (taps * Cin); const int r = idx % (taps * Cin); col = (r / Cin) * Cin4 + (r % Cin); }
  }
  float s = 0.f;
  if (valid) {
    const float* base = partial + (int64_t)co * ncols + col;
#pragma unroll 10
    for (int p = grp; p < nparts; p += 4) s += __ldg(base + (int64_t)p * 64 * ncols);
  }
  s_part[grp][lane_o] = s;
  __syncthreads();
  if (grp == 0 && valid) {
    const float tot = ((s_part[0][lane_o] + s_part[1][lane_o]) + s_part[2][lane_o]) + s_part[3][lane_o];
    if (isb) db[co] += tot;
    else G[idx] = tot;
  }
}

// x (B,T,3) -> [B][1 chunk][T][4] (channel 3 = 0), TF32-rounded
__global__ void pack_x4_kernel(const float* __restrict__ x, float* __restrict__ x4, int64_t n_rows, int C) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = rna_tf32(__ldg(x + i * C));
    if (C > 1) v.y = rna_tf32(__ldg(x + i * C + 1));
    if (C > 2) v.z = rna_tf32(__ldg(x + i * C + 2));
    reinterpret_cast<float4*>(x4)[i] = v;
  }
}

// pooled[b][c*8+bin] = mean_{t in bin} a[b][c/4][t][c%4]     (AdaptiveAvgPool1d(8), models.py:312-315)
__global__ void pool_fwd_chunk_kernel(const float* __restrict__ a, float* __restrict__ pooled, int64_t B, int C, int Tseq) {
  // one thread = (gesture, channel chunk, bin): Tseq / 8 (sixteen for T = 128) contiguous 16-byte loads, four channel means
  const int Cc = C / 4;
  const int W = Tseq / 8;
  const float inv = 1.f / (float)W;
  const int64_t n = B * Cc * 8;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int bin = (int)(i & 7);
    const uint32_t rest = (uint32_t)(i >> 3);          // b * Cc + q
    const uint32_t q = rest % (uint32_t)Cc, b = rest / (uint32_t)Cc;
    const float4* p = a4 + (int64_t)rest * Tseq + bin * W;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16
    for (int t = 0; t < W; ++t) {
      const float4 v = __ldg(p + t);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* o = pooled + (int64_t)b * C * 8 + (q * 4) * 8 + bin;
    o[0] = s.x * inv; o[8] = s.y * inv; o[16] = s.z * inv; o[24] = s.w * inv;
  }
}

// dpre3[b][c/4][t][c%4] = LeakyReLU'(a3) * (dpool[b][c*8 + t/16] / 16 + dfeat)     (un-pool + LeakyReLU backward)
__global__ void unpool_leaky_chunk_kernel(const float* __restrict__ dpool, const float* __restrict__ a3,
                                          const float* __restrict__ dfeat, float* __restrict__ dpre, int64_t B, int C,
                                          int Tseq) {
  // one thread = one 16-byte group (4 channels of one time step)
  const int Cc = C / 4;
  const int W = Tseq / 8;
  const float inv = 1.f / (float)W;
  const int64_t n4 = B * Cc * Tseq;
  const float4* a4 = reinterpret_cast<const float4*>(a3);
  const float4* f4 = reinterpret_cast<const float4*>(dfeat);
  float4* o4 = reinterpret_cast<float4*>(dpre);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % Tseq);
    const uint32_t rest = (uint32_t)(i / Tseq);        // b * Cc + q
    const uint32_t q = rest % (uint32_t)Cc, b = rest / (uint32_t)Cc;
    const float* dp = dpool + (int64_t)b * C * 8 + (q * 4) * 8 + t / W;
    float4 g = make_float4(__ldg(dp) * inv, __ldg(dp + 8) * inv, __ldg(dp + 16) * inv, __ldg(dp + 24) * inv);
    if (dfeat) {
      const float4 f = __ldg(f4 + i);
      g.x += f.x; g.y += f.y; g.z += f.z; g.w += f.w;
    }
    const float4 a = __ldg(a4 + i);
    o4[i] = make_float4(rna_tf32(a.x > 0.f ? g.x : kLeak * g.x), rna_tf32(a.y > 0.f ? g.y : kLeak * g.y),
                        rna_tf32(a.z > 0.f ? g.z : kLeak * g.z), rna_tf32(a.w > 0.f ? g.w : kLeak * g.w));
  }
}

// [B][Cc][T][4] (channel-chunked) -> (B, T, Cc*4) row-major, both sides coalesced through shared memory.
// block = (b, 32 consecutive t); 256 threads
__global__ void __launch_bounds__(256) chunk_to_rows_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                                            int Cc, int Tseq) {
  __shared__ float4 tile[32][17];
  const int64_t b = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < Cc * 32; i += 256) {
    const int q = i / 32, tt = i % 32;
    tile[tt][q] = __ldg(in + (b * Cc + q) * Tseq + t0 + tt);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cc * 32; i += 256) {
    const int tt = i / Cc, q = i % Cc;
    out[(b * Tseq + t0 + tt) * Cc + q] = tile[tt][q];
  }
}

}  // namespace ctc

// ---------------------------------------------------------------------------------------------
// host launchers (used by disc.cu)
// ---------------------------------------------------------------------------------------------
int conv_tc_grid(wgg_ctx* ctx, int64_t items) { return (int)(items < ctx->sm_count ? items : ctx->sm_count); }

// sequence lengths the tcgen05 conv kernels take: whole 128-row tiles
bool conv_tc_seq_ok(int T) { return T >= ctc::T && T % ctc::T == 0 && T <= 1024; }

int conv_tc_fwd_launch(wgg_ctx* ctx, const float* in, const float* wimg, const float* bias, float* out,
                       const float* act_lower, const float* dfeat, int64_t B, int T, int CinC, int taps, int pad, int N,
                       int mode, const char* tag, cudaStream_t st) {
  if (!conv_tc_seq_ok(T) || pad > ctc::PAD_ROWS || taps > 8)
    return wgg_fail(ctx, WGG_EUNSUPPORTED, "conv_tc_fwd: unsupported sequence length / window%s");
  ctc::FwdArgs a;
  a.Tseq = T; a.halves = T / ctc::T;
  a.in = in; a.wimg = wimg; a.bias = bias; a.out = out; a.act_lower = act_lower; a.dfeat = dfeat; a.B = B;
  a.CinC = CinC; a.taps = taps; a.taps_p = (CinC == 1) ? taps + (taps & 1) : taps;
  a.tap_row0 = ctc::PAD_ROWS - pad; a.N = N; a.mode = mode; a.gerr = ctx->async_err;
  const int Kchunks = CinC == 1 ? a.taps_p : taps * CinC;
  const int WCS = (N / 8) * 128;
  const size_t smem = (size_t)((Kchunks * WCS + 1023) / 1024) * 1024 + ctc::NST * (size_t)CinC * ctc::CS + 64 * 4 + (2 * ctc::NST + 2) * 8 + 16;
  const bool multi = a.halves > 1;
  // raise the kernel's dynamic shared memory limit once per context (= per device) to the opt-in maximum: the layers of a
  // model need different amounts
  if (smem > 227 * 1024 || !(multi ? wgg_smem_ok(ctx, ctc::conv_tc_fwd_kernel<true>, 227 * 1024)
                                    : wgg_smem_ok(ctx, ctc::conv_tc_fwd_kernel<false>, 227 * 1024)))
    return wgg_fail(ctx, WGG_ECUDA, "conv_tc_fwd_kernel: cannot reserve shared memory%s");
  const double flops = 2.0 * (double)B * T * N * (double)(taps * CinC * 4);
  ProfScope prof(ctx, "conv_tc_fwd_kernel", st, flops, (double)B * T * 4.0 * (CinC * 4 + N), tag);
  if (multi) ctc::conv_tc_fwd_kernel<true><<<conv_tc_grid(ctx, B * a.halves), ctc::NTHREADS, smem, st>>>(a);
  else ctc::conv_tc_fwd_kernel<false><<<conv_tc_grid(ctx, B), ctc::NTHREADS, smem, st>>>(a);
  WGG_CHECK_LAUNCH(ctx, "conv_tc_fwd_kernel");
  return WGG_OK;
}

int64_t conv_tc_wgrad_ws_floats(wgg_ctx* ctx, int ncols_max) { return (int64_t)ctx->sm_count * 64 * ncols_max; }

// G (Cout x taps*Cin, forward layout) is overwritten; db (Cout) is accumulated.
int conv_tc_wgrad_launch(wgg_ctx* ctx, const float* dpre, const float* in, int64_t B, int T, int Cout, int Cin, int taps,
                         int pad, float* G, float* db, float* ws, cudaStream_t st) {
  const int CoutC = Cout / 4, CinC = (Cin + 3) / 4;
  if (Cout > 64 || CoutC % 8 != 0 || (CinC != 1 && CinC != 16) || taps > 8 || !conv_tc_seq_ok(T) || pad > ctc::W3_XPAD ||
      taps - 1 - pad > ctc::W3_XPAD - 1)
    return wgg_fail(ctx, WGG_EUNSUPPORTED, "conv_tc_wgrad: unsupported layer shape%s");
  ctc::WgradArgs a;
  a.Tseq = T; a.halves = T / ctc::T;
  a.dpre = dpre; a.in = in; a.partial = ws; a.B = B; a.CoutC = CoutC; a.CinC = CinC; a.taps = taps; a.pad = pad;
  const int Cin4 = CinC == 1 ? 4 : 64;
  const int Ktot = CinC == 1 ? 32 : taps * 64;
  a.ncols = Ktot + 8;
  a.gerr = ctx->async_err;
  const bool multi = a.halves > 1;
  const size_t smem = (size_t)ctc::W1_NST * ctc::W1_STAGE + 1024 + 16 * 8 + 16;
  if (!wgg_smem_ok(ctx, ctc::conv_tc_wgrad_kernel<false>, smem) || !wgg_smem_ok(ctx, ctc::conv_tc_wgrad_kernel<true>, smem))
    return wgg_fail(ctx, WGG_ECUDA, "conv_tc_wgrad_kernel: cannot reserve shared memory%s");
  const int grid = conv_tc_grid(ctx, B * a.halves);
  {
    ProfScope prof(ctx, "conv_tc_wgrad_kernel", st, 2.0 * (double)B * T * Cout * (double)(taps * Cin),
                   (double)B * T * 4.0 * (Cout + CinC * 4), "conv_tc_wgrad_kernel");
    if (CinC == 16) {
      const size_t smem3 = (size_t)2 * ctc::W3_STAGE + 2 * (multi ? ctc::W3Geo<true>::RAW_SLOT : ctc::W3Geo<false>::RAW_SLOT) +
                           16 * 64 * 4 + 10 * 8 + 16;
      if (!(multi ? wgg_smem_ok(ctx, ctc::conv_tc_wgrad3_kernel<true>, smem3) : wgg_smem_ok(ctx, ctc::conv_tc_wgrad3_kernel<false>, smem3)))
        return wgg_fail(ctx, WGG_ECUDA, "conv_tc_wgrad3_kernel: cannot reserve shared memory%s");
      if (multi) ctc::conv_tc_wgrad3_kernel<true><<<grid, ctc::W3_THREADS, smem3, st>>>(a);
      else ctc::conv_tc_wgrad3_kernel<false><<<grid, ctc::W3_THREADS, smem3, st>>>(a);
    } else if (multi) {
      ctc::conv_tc_wgrad_kernel<true><<<grid, ctc::W_THREADS, smem, st>>>(a);
    } else {
      ctc::conv_tc_wgrad_kernel<false><<<grid, ctc::W_THREADS, smem, st>>>(a);
    }
    WGG_CHECK_LAUNCH(ctx, "conv_tc_wgrad_kernel");
  }
  ctc::wgrad_finalize_kernel<<<(Cout * taps * Cin + Cout + 63) / 64, 256, 0, st>>>(ws, grid, a.ncols, Cout, Cin, Cin4, taps, Ktot, G, db);
  WGG_CHECK_LAUNCH(ctx, "wgrad_finalize_kernel");
  return WGG_OK;
}

int pack_x4_launch(wgg_ctx* ctx, const float* x, float* x4, int64_t B, int T, int C, cudaStream_t st) {
  ctc::pack_x4_kernel<<<ew_blocks(B * T), 256, 0, st>>>(x, x4, B * T, C);
  WGG_CHECK_LAUNCH(ctx, "pack_x4_kernel");
  return WGG_OK;
}

int pool_fwd_chunk_launch(wgg_ctx* ctx, const float* a, float* pooled, int64_t B, int T, int C, cudaStream_t st) {
  ctc::pool_fwd_chunk_kernel<<<ew_blocks(B * (C / 4) * 8), 256, 0, st>>>(a, pooled, B, C, T);
  WGG_CHECK_LAUNCH(ctx, "pool_fwd_chunk_kernel");
  return WGG_OK;
}

int unpool_leaky_chunk_launch(wgg_ctx* ctx, const float* dpool, const float* a3, const float* dfeat, float* dpre,
                              int64_t B, int T, int C, cudaStream_t st) {
  ctc::unpool_leaky_chunk_kernel<<<ew_blocks(B * (C / 4) * T), 256, 0, st>>>(dpool, a3, dfeat, dpre, B, C, T);
  WGG_CHECK_LAUNCH(ctx, "unpool_leaky_chunk_kernel");
  return WGG_OK;
}

// ---- debug / test hooks (exported): run one tcgen05 conv kernel on caller-provided chunk-layout tensors ----
extern "C" __attribute__((visibility("default"))) int wgg_debug_conv_tc_wgrad(wgg_ctx* ctx, const float* dpre,
                                                                              const float* in, int64_t B, int Cout,
                                                                              int Cin, int taps, int pad, float* G,
                                                                              float* db, float* ws, void* stream) {
  if (!ctx) return WGG_EINVAL;
  return conv_tc_wgrad_launch(ctx, dpre, in, B, ctc::T, Cout, Cin, taps, pad, G, db, ws, (cudaStream_t)stream);
}
extern "C" __attribute__((visibility("default"))) int wgg_debug_conv_tc_fwd(wgg_ctx* ctx, const float* in,
                                                                            const float* wimg, const float* bias,
                                                                            float* out, int64_t B, int CinC, int taps,
                                                                            int pad, int N, int mode, void* stream) {
  if (!ctx) return WGG_EINVAL;
  return conv_tc_fwd_launch(ctx, in, wimg, bias, out, nullptr, nullptr, B, ctc::T, CinC, taps, pad, N, mode, "debug",
                            (cudaStream_t)stream);
}

int chunk_to_rows_launch(wgg_ctx* ctx, const float* in, float* out, int64_t B, int T, int C, cudaStream_t st) {
  if (C % 4 != 0 || C / 4 > 16 || B > 65535 || T % 32) return wgg_fail(ctx, WGG_EINVAL, "chunk_to_rows: unsupported shape%s");
  dim3 grid(T / 32, (unsigned)B);
  ctc::chunk_to_rows_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), C / 4, T);
  WGG_CHECK_LAUNCH(ctx, "chunk_to_rows_kernel");
  return WGG_OK;
}


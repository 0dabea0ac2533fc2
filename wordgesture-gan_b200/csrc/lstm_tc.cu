// Persistent BiLSTM-layer forward on 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces nn.LSTM's per-layer work of src/gan/models.py:160 for every generator call of the step (the no-grad
// critic-phase calls, src/shared/utils.py:70-73,91-94; the grad-carrying calls of the G/E step, which also stash what
// BPTT needs; and all of sampling, eval_gan.py:131-135):
//     gates_t = x_t W_ih^T + h_{t-1} W_hh^T + b ;  c_t = f c_{t-1} + i g ;  h_t = o tanh(c_t)
// One CTA owns a tile of 128 samples of one direction for all T timesteps:
//   * W_ih, W_hh (TF32) stay resident in shared memory for the whole sequence;
//   * x_t tiles stream in through a 5-stage ring of bulk asynchronous copies (cp.async.bulk -> mbarrier tx);
//   * one warp issues tcgen05.mma kind::tf32 (M=128 samples, N=192 gates, K=8 per instruction; warp-uniform loop,
//     one elected lane), the accumulator lives in TMEM (two 192-column buffers: the x-projection of step t+1 is
//     issued while the epilogue of step t still runs; h_{t-1} W_hh^T follows in three phases as the h chunks land);
//   * 16 epilogue warps read the accumulator with tcgen05.ld (one thread = one sample row and 12 hidden units,
//     gates of a hidden unit are adjacent columns because the weight rows are permuted unit-major), apply the gate
//     non-linearities, keep c in registers, and write h_t both to shared memory (the next step's A operand)
//     and to HBM (the next layer's input).
// The rest of this file is the backward of the stack on the same layouts: BPTT, input gradient, weight gradient,
// and the fused output-head backward.
//
// Operand layout ("tc layout", no swizzle, K-major core matrices): a [rows x K] fp32 matrix is stored as
//   [K/4 chunks][rows/8 groups][8 rows][4 floats]   i.e. 128-byte core matrices (8 rows x 16 B);
// UMMA descriptors use LBO = stride between consecutive 16-byte K chunks, SBO = 128 B between 8-row groups.
// Layer activations live in HBM in exactly this layout per 128-row tile and timestep, so a timestep's tile is
// ONE contiguous block: it is fetched with plain bulk copies (no tensor map) and written by the epilogue with
// fully coalesced 16-byte stores (lane = row).
#include "tc_common.cuh"

namespace tc {

constexpr int TM = 128;   // samples per CTA tile (UMMA M)
constexpr int HID = 48;   // hidden units (this kernel is specialised for the default model)
constexpr int N4 = 192;   // gate columns (UMMA N)
constexpr int KH_CHUNKS = HID / 4;
constexpr int CHUNK_BYTES_A = (TM / 8) * 128;  // 2048: one 16-byte K chunk of a 128-row tile
constexpr int CHUNK_BYTES_W = (N4 / 8) * 128;  // 3072: one 16-byte K chunk of the 192-row weight image
constexpr int SB_CHUNKS = 8;                   // K chunks per ring stage (32 floats)
constexpr int SB_BYTES = SB_CHUNKS * CHUNK_BYTES_A;  // 16384
constexpr int NSTAGE = 5;
// warp 0 producer, warp 1 MMA issuer, then 4 * (12 / NCH) epilogue warps: every epilogue thread owns one gesture row and
// NCH K-chunks (4 * NCH hidden units) of it.  NCH = 3: 16 epilogue warps (576 threads) is what ships; NCH = 2 (24 epilogue
// warps, 6 per scheduler) was measured SLOWER on B200 (6.91 vs 6.47 ms for a 40 960-gesture call): the gate math is
// bound by the MUFU / MIO instruction stream, not by latency hiding.
__host__ __device__ constexpr int fwd_epi_warps(int nch) { return 4 * (KH_CHUNKS / nch); }
__host__ __device__ constexpr int fwd_threads(int nch) { return 64 + 32 * fwd_epi_warps(nch); }
constexpr int ACC_COLS = N4;                   // TMEM columns per accumulator buffer
constexpr int TMEM_COLS = 512;

using namespace tcu;

// instruction descriptor: D=f32, A=B=tf32, both K-major, N=192, M=128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N4 >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// element (row, k) of a tc-layout tile with `rows` rows: float index
__host__ __device__ __forceinline__ int64_t tc_index(int rows, int row, int k) {
  return ((int64_t)(k >> 2) * (rows >> 3) + (row >> 3)) * 32 + (row & 7) * 4 + (k & 3);
}

// ---------------------------------------------------------------------------------------------
// weight image: per (layer, dir): Wx [KXC chunks][24][8][4], Wh [12][24][8][4], bias[192]; rows permuted
// unit-major (n' = 4u + gate  <->  PyTorch row gate*H + u), values rounded to TF32 (bias stays fp32).
// ---------------------------------------------------------------------------------------------
__global__ void prep_weights_kernel(const float* __restrict__ lp, int64_t dir_stride, int64_t off_whh, int64_t off_bih,
                                    int64_t off_bhh, int I, int KX, float* __restrict__ img, int64_t img_stride) {
  const int dir = blockIdx.y;
  const float* w = lp + dir * dir_stride;
  float* o = img + dir * img_stride;
  const int nx = N4 * KX, nh = N4 * HID;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nx + nh + N4; idx += gridDim.x * blockDim.x) {
    if (idx < nx) {
      const int np = idx / KX, k = idx % KX;
      const int u = np >> 2, g = np & 3;
      const float v = k < I ? w[(int64_t)(g * HID + u) * I + k] : 0.f;
      o[tc_index(N4, np, k)] = rna_tf32(v);
    } else if (idx < nx + nh) {
      const int j = idx - nx;
      const int np = j / HID, k = j % HID;
      const int u = np >> 2, g = np & 3;
      o[nx + tc_index(N4, np, k)] = rna_tf32(w[off_whh + (int64_t)(g * HID + u) * HID + k]);
    } else {
      const int np = idx - nx - nh;
      const int u = np >> 2, g = np & 3;
      o[nx + nh + np] = w[off_bih + g * HID + u] + w[off_bhh + g * HID + u];
    }
  }
}

// layer-0 input in tc layout, K padded to KX0 (zero fill), rows padded to the tile: (models.py:147-157)
__global__ void __launch_bounds__(256) build_x0_tc_kernel(const float* __restrict__ proto, const float* __restrict__ z,
                                                          float* __restrict__ x0, int T, int64_t B, int ntiles, int C,
                                                          int pd, int Z, int KX0) {
  // block = (tile, 8 timesteps): the latent part of x0 is the same for every timestep, so the tile's z rows are
  // staged once in shared memory (coalesced) and re-used; every store is a coalesced 16-byte group.
  extern __shared__ float s_z[];  // [TM][Z + 1]
  const int tile = blockIdx.x, t0 = blockIdx.y * 8;
  const int KC = KX0 / 4, ZP = Z + 1;
  for (int i = threadIdx.x; i < TM * Z; i += 256) {
    const int row = i / Z, j = i % Z;
    const int64_t b = (int64_t)tile * TM + row;
    s_z[row * ZP + j] = b < B ? __ldg(z + b * Z + j) : 0.f;
  }
  __syncthreads();
  float4* o4 = reinterpret_cast<float4*>(x0);
  for (int tt = 0; tt < 8 && t0 + tt < T; ++tt) {
    const int t = t0 + tt;
    for (int i = threadIdx.x; i < KC * TM; i += 256) {
      const int chunk = i >> 7, row = i & (TM - 1);
      const int64_t b = (int64_t)tile * TM + row;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (b < B) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = chunk * 4 + j;
          if (k < pd) v[j] = __ldg(proto + (b * T + t) * C + k);
          else if (k < pd + Z) v[j] = s_z[row * ZP + (k - pd)];
        }
      }
      o4[((int64_t)t * ntiles + tile) * KC * TM + i] = make_float4(rna_tf32(v[0]), rna_tf32(v[1]), rna_tf32(v[2]), rna_tf32(v[3]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the persistent layer kernel.  grid (ntiles, 2 directions), 576 threads, 1 CTA / SM.
// xin : [T][ntiles][KXC][16][8][4]   hout : [T][ntiles][24][16][8][4] (this direction fills chunks dir*12..+12)
// ---------------------------------------------------------------------------------------------
// STASH = 1 (grad-carrying forward) additionally writes, per step, what BPTT needs:
//   gc   : [2 dirs][T][ntiles][GC_CHUNKS][128 rows][4]  chunk u < 48 = (i,f,g,o) of hidden unit u, chunk 48 + u/4 = c
//   h_rm : [T][B][2H] row-major, un-rounded fp32 (operand of the weight-gradient GEMMs and of the output head)
constexpr int GC_CHUNKS = HID + HID / 4;  // 60

template <int KXC, int STASH, int NCH>
__global__ void __launch_bounds__(fwd_threads(NCH), 1) lstm_tc_fwd_kernel(const float* __restrict__ xin,
                                                                  const float* __restrict__ wimg, int64_t img_stride,
                                                                  float* __restrict__ hout, int T, int ntiles,
                                                                  float* __restrict__ gc, float* __restrict__ h_rm,
                                                                  int64_t B, int* __restrict__ gerr) {
  constexpr int NSB = (KXC + SB_CHUNKS - 1) / SB_CHUNKS;
  constexpr int WX_BYTES = KXC * CHUNK_BYTES_W;
  constexpr int WH_BYTES = KH_CHUNKS * CHUNK_BYTES_W;
  constexpr int NTHREADS = fwd_threads(NCH), EPI_WARPS = fwd_epi_warps(NCH);
  constexpr int NPART = KH_CHUNKS / NCH;  // column parts of the accumulator (one per group of four epilogue warps)
  static_assert(KH_CHUNKS % NCH == 0 && NPART % 2 == 0, "h chunks are handed over in NCH phases of NPART / 2 MMAs");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_wx = smem;
  uint8_t* s_wh = s_wx + WX_BYTES;
  uint8_t* s_x = s_wh + WH_BYTES;
  uint8_t* s_h = s_x + NSTAGE * SB_BYTES;
  float* s_bias = reinterpret_cast<float*>(s_h + KH_CHUNKS * CHUNK_BYTES_A);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + N4);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 20);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, dir = blockIdx.y;
  const uint32_t bar0 = smem_u32(s_bar);
  auto BAR_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto BAR_ACC_FULL = [&](int b) { return bar0 + 8u * (2 * NSTAGE + b); };
  auto BAR_ACC_EMPTY = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 2 + b); };
  auto BAR_H = [&](int ph) { return bar0 + 8u * (2 * NSTAGE + 4 + ph); };  // h chunks {ph + NCH j} written

  // ---- one-time setup -------------------------------------------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(wimg + dir * img_stride);
    float4* dst = reinterpret_cast<float4*>(s_wx);
    constexpr int n4 = (WX_BYTES + WH_BYTES) / 16;
    for (int i = tid; i < n4; i += NTHREADS) dst[i] = __ldg(src + i);
    const float* bsrc = wimg + dir * img_stride + (WX_BYTES + WH_BYTES) / 4;
    // bias pre-scaled for the exponent form of the gates: -b log2(e) for i, f, o and -2 b log2(e) for g
    for (int i = tid; i < N4; i += NTHREADS) s_bias[i] = __ldg(bsrc + i) * ((i & 3) == 2 ? -2.f * kLog2e : -kLog2e);
  }
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(BAR_FULL(s), 1);
      mbar_init(BAR_EMPTY(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(BAR_ACC_FULL(b), 1);
      mbar_init(BAR_ACC_EMPTY(b), EPI_WARPS);
    }
    for (int ph = 0; ph < NCH; ++ph) mbar_init(BAR_H(ph), EPI_WARPS);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();  // weights were written through the generic proxy; the tensor core reads via the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== producer: stream x_t sub-blocks =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int step = 0; step < T && ok; ++step) {
        const int t = dir ? T - 1 - step : step;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(xin) + ((int64_t)t * ntiles + tile) * KXC * CHUNK_BYTES_A;
#pragma unroll 1
        for (int sb = 0; sb < NSB; ++sb) {
          if (!mbar_wait(BAR_EMPTY(stage), phase ^ 1, s_abort, gerr, 1)) { ok = false; break; }
          const int chunks = (KXC - sb * SB_CHUNKS) < SB_CHUNKS ? (KXC - sb * SB_CHUNKS) : SB_CHUNKS;
          const uint32_t bytes = chunks * CHUNK_BYTES_A;
          mbar_expect_tx(BAR_FULL(stage), bytes);
          bulk_g2s(smem_u32(s_x + stage * SB_BYTES), src + (int64_t)sb * SB_BYTES, bytes, BAR_FULL(stage));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues (operands stay in uniform registers) =====
    int stage = 0;
    uint32_t phase = 0;
    bool ok = true;
    const uint32_t wx = smem_u32(s_wx), wh = smem_u32(s_wh), hs = smem_u32(s_h);
    const uint64_t bdx0 = make_desc(wx, CHUNK_BYTES_W, 128);
    const uint64_t adh3 = make_desc(hs, NCH * CHUNK_BYTES_A, 128), bdh3 = make_desc(wh, NCH * CHUNK_BYTES_W, 128);  // chunk pairs (c, c+NCH)
    constexpr uint64_t kAStep = (2 * CHUNK_BYTES_A) >> 4, kWStep = (2 * CHUNK_BYTES_W) >> 4;
    for (int step = 0; step < T && ok; ++step) {
      const int b = step & 1;
      const uint32_t use = (uint32_t)(step >> 1);  // how many times this buffer has been used before
      if (!mbar_wait(BAR_ACC_EMPTY(b), (use & 1) ^ 1, s_abort, gerr, 2)) break;
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(b * ACC_COLS);
#pragma unroll 1
      for (int sb = 0; sb < NSB; ++sb) {
        if (!mbar_wait(BAR_FULL(stage), phase, s_abort, gerr, 3)) { ok = false; break; }
        tc_fence_after();
        const int chunks = (KXC - sb * SB_CHUNKS) < SB_CHUNKS ? (KXC - sb * SB_CHUNKS) : SB_CHUNKS;
        if (elect_one()) {
          uint64_t ad = make_desc(smem_u32(s_x + stage * SB_BYTES), CHUNK_BYTES_A, 128);
          uint64_t bd = bdx0 + (uint64_t)(sb * SB_CHUNKS / 2) * kWStep;
          for (int j = 0; j < chunks / 2; ++j, ad += kAStep, bd += kWStep) mma_tf32_ss(tacc, ad, bd, kIdesc, (sb | j) ? 1u : 0u);
          mma_commit(BAR_EMPTY(stage));  // frees the ring stage once these MMAs have read it
        }
        __syncwarp();
        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (!ok) break;
      // recurrent part, pipelined against the gate math of the previous step: every epilogue warp writes its h chunks
      // in NCH phases; as soon as phase ph is complete the MMAs over the K-chunk pairs (ph + 2 NCH g, ph + 2 NCH g + NCH)
      // are issued (the pair's second chunk is addressed through the descriptor's leading-byte-offset), so only the
      // last 1 / NCH of W_hh h is exposed after the gate math
      if (step > 0) {
        for (int ph = 0; ph < NCH; ++ph) {
          if (!mbar_wait(BAR_H(ph), (uint32_t)((step - 1) & 1), s_abort, gerr, 4)) { ok = false; break; }
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int g2 = 0; g2 < NPART / 2; ++g2)
              mma_tf32_ss(tacc, adh3 + (uint64_t)(((2 * NCH * g2 + ph) * CHUNK_BYTES_A) >> 4),
                          bdh3 + (uint64_t)(((2 * NCH * g2 + ph) * CHUNK_BYTES_W) >> 4), kIdesc, 1u);
          }
          __syncwarp();
        }
        if (!ok) break;
      }
      if (elect_one()) mma_commit(BAR_ACC_FULL(b));
      __syncwarp();
    }
  } else {
    // ===== epilogue: EPI_WARPS warps; TMEM lane quarter = warp % 4, column part = (warp - 2) / 4 (UPT hidden units each) =====
    // Four or six warps per scheduler keep the MUFU pipe (5 EX2 + 2 RCP per cell) and the FMA pipe busy at the same time.
    constexpr int UPT = 4 * NCH;  // hidden units per thread
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    float c[UPT];
#pragma unroll
    for (int i = 0; i < UPT; ++i) c[i] = 0.f;
    float4* hs4 = reinterpret_cast<float4*>(s_h);
    for (int step = 0; step < T; ++step) {
      const int t = dir ? T - 1 - step : step;
      const int b = step & 1;
      if (!mbar_wait(BAR_ACC_FULL(b), (uint32_t)((step >> 1) & 1), s_abort, gerr, 5)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * ACC_COLS + part * 4 * UPT);
      float4* hg4 = reinterpret_cast<float4*>(hout) +
                    (((int64_t)t * ntiles + tile) * (2 * KH_CHUNKS) + dir * KH_CHUNKS) * (TM) + row;
      float4* gc4 = nullptr;
      if (STASH)
        gc4 = reinterpret_cast<float4*>(gc) + ((((int64_t)dir * T + t) * ntiles + tile) * GC_CHUNKS) * TM + row;
      float v[16 * NCH];
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        float vq[16];
        tmem_ld16(taddr + 16 * q, vq);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[16 * q + i] = vq[i];
      }
      // accumulator fully read by this warp: hand the buffer back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_ACC_EMPTY(b));
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int chunk = part * NCH + ch;            // K chunk (4 hidden units) of this direction's h
        float hv[4], hraw[4];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int ul = ch * 4 + uu;                 // unit index inside this thread's UPT
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + (part * UPT + ul) * 4);
          // 5 gate non-linearities with 5 EX2 + 2 RCP (instead of 5 + 5): the sigmoids / tanh of a unit share their
          // reciprocals.  sigma(x) = 1/(1+E), E = 2^(-x log2 e); tanh(x) = (1-E2)/(1+E2), E2 = 2^(-2x log2 e).  The
          // scale and the (pre-scaled) bias are one FMA; exponents are capped at 40 so that the product of three
          // (1+E) stays finite (2^-40 = 9e-13 is below fp32 resolution of the gate values).
          const float Ei = ex2_fast(fminf(fmaf(v[4 * ul + 0], -kLog2e, bb.x), 40.f));
          const float Ef = ex2_fast(fminf(fmaf(v[4 * ul + 1], -kLog2e, bb.y), 40.f));
          const float Eg = ex2_fast(fminf(fmaf(v[4 * ul + 2], -2.f * kLog2e, bb.z), 40.f));
          const float Eo = ex2_fast(fminf(fmaf(v[4 * ul + 3], -kLog2e, bb.w), 40.f));
          const float pi = 1.f + Ei, pf = 1.f + Ef, pg = 1.f + Eg, po = 1.f + Eo;
          const float pfpi = pf * pi;
          const float r1 = rcp_fast(pfpi * pg);
          const float tg = r1 * pg;
          const float ig = tg * pf;
          const float fg = tg * pi;
          const float qg = r1 * pfpi;                 // 1 / (1 + Eg)
          const float gg = fmaf(-Eg, qg, qg);         // tanh
          c[ul] = fmaf(fg, c[ul], ig * gg);
          const float Ec = ex2_fast(fminf(c[ul] * (-2.f * kLog2e), 40.f));
          const float pc = 1.f + Ec;
          const float r2 = rcp_fast(po * pc);
          const float og = r2 * pc;
          hraw[uu] = fmaf(-Ec, r2, r2);               // = og * tanh(c)
          hv[uu] = rna_tf32(hraw[uu]);
          if (STASH) gc4[(int64_t)(part * UPT + ul) * TM] = make_float4(ig, fg, gg, og);
        }
        const float4 q = make_float4(hv[0], hv[1], hv[2], hv[3]);
        hs4[chunk * TM + row] = q;                    // next step's A operand
        hg4[(int64_t)chunk * TM] = q;                 // next layer's input (coalesced: lane = row)
        if (STASH) {
          gc4[(int64_t)(HID + chunk) * TM] = make_float4(c[ch * 4], c[ch * 4 + 1], c[ch * 4 + 2], c[ch * 4 + 3]);
          const int64_t bidx = (int64_t)tile * TM + row;
          if (h_rm != nullptr && bidx < B)
            *reinterpret_cast<float4*>(h_rm + ((int64_t)t * B + bidx) * (2 * HID) + dir * HID + chunk * 4) =
                make_float4(hraw[0], hraw[1], hraw[2], hraw[3]);
        }
        fence_async_smem();   // this phase's h chunk is visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR_H(ch));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// BPTT on tensor cores.  One CTA = 128 samples of one direction, reverse scan over time.  Per step the epilogue
// warps turn (gates, c, c_prev, dh) into d(pre-activation) `da` [128 x 192] (written to shared memory as the next
// MMA's A operand and to HBM row-major for the weight-gradient / input-gradient GEMMs), and one elected thread
// issues dh_rec = da * W_hh (M=128, N=48, K=192) whose accumulator the NEXT (earlier) step reads from TMEM.
//   gc     : [2][T][ntiles][60][128][4]   (forward stash)
//   dh_out : [T][ntiles][24][128][4]      d(layer output), chunk layout (chunks dir*12.. belong to this direction)
//   da_out : [2][T][ntiles][48][128][4]   chunk u = (da_i, da_f, da_g, da_o) of hidden unit u - exactly the A tile
// ---------------------------------------------------------------------------------------------
constexpr int BWD_THREADS = 256;  // 8 epilogue warps (TMEM quarter = warp % 4); warp 0 also issues the MMAs
constexpr int WT_CHUNK_BYTES = (HID / 8) * 128;  // 768: one K chunk of the [48 x 192] W_hh^T image

// B operand of dh_rec = da * W_hh:  img[u'][n'] = W_hh[(g*H + u)][u'] with n' = 4u + g   ([N=48][K=192], K-major)
__global__ void prep_whhT_kernel(const float* __restrict__ lp, int64_t dir_stride, int64_t off_whh,
                                 float* __restrict__ img) {
  const int dir = blockIdx.y;
  const float* w = lp + dir * dir_stride + off_whh;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < HID * N4; idx += gridDim.x * blockDim.x) {
    const int up = idx / N4, kk = idx % N4;
    const int u = kk >> 2, g = kk & 3;
    img[dir * HID * N4 + tc_index(HID, up, kk)] = rna_tf32(w[(int64_t)(g * HID + u) * HID + up]);
  }
}

__global__ void __launch_bounds__(BWD_THREADS, 1) lstm_tc_bwd_kernel(const float* __restrict__ gc,
                                                                     const float* __restrict__ wimg,
                                                                     const float* __restrict__ dh_out,
                                                                     float* __restrict__ da_out, int T, int ntiles,
                                                                     int64_t B, int* __restrict__ gerr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_da = smem;                                   // [48 chunks][128 rows][16 B]
  uint8_t* s_w = s_da + HID * CHUNK_BYTES_A;              // [48 chunks][6 groups][128 B]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_w + HID * WT_CHUNK_BYTES);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 4);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, dir = blockIdx.y;
  const uint32_t BAR_DA = smem_u32(s_bar), BAR_ACC = smem_u32(s_bar) + 8;
  {
    const float4* src = reinterpret_cast<const float4*>(wimg + dir * HID * N4);
    float4* dst = reinterpret_cast<float4*>(s_w);
    for (int i = tid; i < HID * N4 / 4; i += BWD_THREADS) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    mbar_init(BAR_DA, 8);
    mbar_init(BAR_ACC, 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 64);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  {
    const uint32_t idesc = make_idesc(TM, HID);
    const uint64_t ad0 = make_desc(smem_u32(s_da), CHUNK_BYTES_A, 128), bd0 = make_desc(smem_u32(s_w), WT_CHUNK_BYTES, 128);
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int row = quarter * 32 + lane;
    const int64_t bidx = (int64_t)tile * TM + row;
    const bool valid = bidx < B;
    float dc[24], cc[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) dc[i] = 0.f;
    float4* da4 = reinterpret_cast<float4*>(s_da);
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 24);
    const float4* gc4 = reinterpret_cast<const float4*>(gc) + row;
    auto gc_tile = [&](int tt) { return gc4 + ((((int64_t)dir * T + tt) * ntiles + tile) * GC_CHUNKS) * TM; };
    {
      // cell state of the first step of the reverse scan; afterwards it is carried over from the c_prev loads
      const float4* g0 = gc_tile(dir ? 0 : T - 1);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float4 c4 = __ldg(g0 + (int64_t)(HID + half * 6 + k) * TM);
        cc[4 * k] = c4.x; cc[4 * k + 1] = c4.y; cc[4 * k + 2] = c4.z; cc[4 * k + 3] = c4.w;
      }
    }
    int n = 0;
    for (int step = T - 1; step >= 0; --step, ++n) {
      const int t = dir ? T - 1 - step : step;
      const int tp = dir ? t + 1 : t - 1;
      const float4* g4 = gc_tile(t);
      const float4* dh4 = reinterpret_cast<const float4*>(dh_out) + (((int64_t)t * ntiles + tile) * (2 * KH_CHUNKS) + dir * KH_CHUNKS + half * 6) * TM + row;
      float4* dao4 = reinterpret_cast<float4*>(da_out) + ((((int64_t)dir * T + t) * ntiles + tile) * HID + half * 24) * TM + row;
      // Every global operand of this step is independent of the recurrence: issue all loads BEFORE waiting for the
      // dh_rec accumulator, so the memory round trip overlaps the tensor-core latency of the previous step.
      float4 gts[24];
      float dho[24], cp[24];
#pragma unroll
      for (int u = 0; u < 24; ++u) gts[u] = __ldg(g4 + (int64_t)(half * 24 + u) * TM);  // (i, f, g, o)
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 a0 = valid ? __ldg(dh4 + (int64_t)k * TM) : z4;
        dho[4 * k] = a0.x; dho[4 * k + 1] = a0.y; dho[4 * k + 2] = a0.z; dho[4 * k + 3] = a0.w;
        float4 p0 = z4;
        if (step > 0) p0 = __ldg(gc_tile(tp) + (int64_t)(HID + half * 6 + k) * TM);
        cp[4 * k] = p0.x; cp[4 * k + 1] = p0.y; cp[4 * k + 2] = p0.z; cp[4 * k + 3] = p0.w;
      }
      if (step < T - 1) {
        if (!mbar_wait(BAR_ACC, (uint32_t)((n - 1) & 1), s_abort, gerr, 32)) break;
        tc_fence_after();
      }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float rec[8];
        if (step < T - 1) tmem_ld8(taddr + ch * 8, rec);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) rec[i] = 0.f;
        }
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          const int ul = ch * 8 + uu;
          const float4 gt = gts[ul];
          const float tch = tanh_fast(cc[ul]);
          const float dh = dho[ul] + rec[uu];
          const float d_o = dh * tch;
          const float dct = dc[ul] + dh * gt.w * (1.f - tch * tch);
          const float dai = dct * gt.z * gt.x * (1.f - gt.x);
          const float daf = dct * cp[ul] * gt.y * (1.f - gt.y);
          const float dag = dct * gt.x * (1.f - gt.z * gt.z);
          const float dao = d_o * gt.w * (1.f - gt.w);
          dc[ul] = dct * gt.y;
          cc[ul] = cp[ul];
          const float4 dq = make_float4(rna_tf32(dai), rna_tf32(daf), rna_tf32(dag), rna_tf32(dao));
          da4[(half * 24 + ul) * TM + row] = dq;       // A operand of dh_rec = da * W_hh
          dao4[(int64_t)ul * TM] = dq;                 // HBM copy for the dx / dW kernels (coalesced: lane = row)
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_DA);
      if (warp == 0 && step >= 1) {
        // dh_rec = da * W_hh for the next (earlier) step: warp-uniform, one elected lane issues
        if (!mbar_wait(BAR_DA, (uint32_t)(n & 1), s_abort, gerr, 31)) break;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < N4 / 8; ++j)
            mma_tf32_ss(tmem_base, ad0 + (uint64_t)(j * ((2 * CHUNK_BYTES_A) >> 4)), bd0 + (uint64_t)(j * ((2 * WT_CHUNK_BYTES) >> 4)),
                        idesc, j ? 1u : 0u);
          mma_commit(BAR_ACC);
        }
        __syncwarp();
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
// dx = sum_dir da_dir * W_ih_dir : the gradient w.r.t. a layer's input (= d(output) of the layer below).
// One CTA = one half of the input features (N = 48) looping over (t, tile) pairs; per pair and direction the
// 96 KB da tile arrives with ONE bulk copy (it is contiguous in the chunk layout) and feeds 24 K-major MMAs
// (M=128 samples, N=48, K=192 gates); both directions accumulate into the same TMEM tile.
//   da   : [2][T][ntiles][48][128][4]       wimg : [2 halves][2 dirs][48 x 192] K-major images of W_ih^T
//   dx   : [T][ntiles][nchunks_out][128][4]  (this CTA writes chunks half*12 .. half*12+11)
// ---------------------------------------------------------------------------------------------
constexpr int DX_THREADS = 192;  // warp 0 producer, warp 1 MMA, warps 2..5 epilogue

// img[half][dir][feat][n'] = W_ih[dir][(g*H + u)][half*48 + feat]  (0 beyond I), n' = 4u + g
__global__ void prep_wihT_kernel(const float* __restrict__ lp, int64_t dir_stride, int I, float* __restrict__ img) {
  const int dir = blockIdx.y, half = blockIdx.z;
  const float* w = lp + dir * dir_stride;
  float* o = img + ((int64_t)half * 2 + dir) * HID * N4;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < HID * N4; idx += gridDim.x * blockDim.x) {
    const int f = idx / N4, kk = idx % N4;
    const int u = kk >> 2, g = kk & 3;
    const int feat = half * HID + f;
    o[tc_index(HID, f, kk)] = feat < I ? rna_tf32(w[(int64_t)(g * HID + u) * I + feat]) : 0.f;
  }
}

__global__ void __launch_bounds__(DX_THREADS, 1) lstm_tc_dx_kernel(const float* __restrict__ da,
                                                                   const float* __restrict__ wimg,
                                                                   float* __restrict__ dx, int T, int ntiles,
                                                                   int out_chunks, int* __restrict__ gerr) {
  constexpr int NSTG = 3;                                 // ring of half da tiles (24 K-chunks = 48 KB each)
  constexpr int HALF_CHUNKS = HID / 2, HALF_BYTES = HALF_CHUNKS * CHUNK_BYTES_A;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_a = smem;
  uint8_t* s_w = s_a + NSTG * HALF_BYTES;               // 2 dirs x (48 chunks x 768 B)
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_w + 2 * HID * WT_CHUNK_BYTES);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * NSTG + 2);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.y;
  const uint32_t bar0 = smem_u32(s_bar);
  auto BAR_FULL = [&](int st) { return bar0 + 8u * st; };
  auto BAR_EMPTY = [&](int st) { return bar0 + 8u * (NSTG + st); };
  const uint32_t BAR_ACC_FULL = bar0 + 8u * (2 * NSTG), BAR_ACC_EMPTY = bar0 + 8u * (2 * NSTG + 1);
  {
    const float4* src = reinterpret_cast<const float4*>(wimg + (int64_t)half * 2 * HID * N4);
    float4* dst = reinterpret_cast<float4*>(s_w);
    for (int i = tid; i < 2 * HID * N4 / 4; i += DX_THREADS) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    for (int st = 0; st < NSTG; ++st) { mbar_init(BAR_FULL(st), 1); mbar_init(BAR_EMPTY(st), 1); }
    mbar_init(BAR_ACC_FULL, 1);
    mbar_init(BAR_ACC_EMPTY, 4);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 64);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int64_t npairs = (int64_t)T * ntiles;
  const int64_t tile_floats = (int64_t)HID * TM * 4;  // one da tile

  if (warp == 0) {
    if (lane == 0) {
      int n = 0;
      for (int64_t pr = blockIdx.x; pr < npairs; pr += gridDim.x)
        for (int dh = 0; dh < 4; ++dh, ++n) {   // (direction, K half)
          const int st = n % NSTG;
          if (!mbar_wait(BAR_EMPTY(st), (uint32_t)(((n / NSTG) & 1) ^ 1), s_abort, gerr, 41)) return;
          mbar_expect_tx(BAR_FULL(st), HALF_BYTES);
          bulk_g2s(smem_u32(s_a) + st * HALF_BYTES,
                   reinterpret_cast<const uint8_t*>(da + ((int64_t)(dh >> 1) * npairs + pr) * tile_floats) + (dh & 1) * HALF_BYTES,
                   HALF_BYTES, BAR_FULL(st));
        }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues
    const uint32_t idesc = make_idesc(TM, HID);
    const uint64_t ad0 = make_desc(smem_u32(s_a), CHUNK_BYTES_A, 128), bd0 = make_desc(smem_u32(s_w), WT_CHUNK_BYTES, 128);
    int n = 0, np = 0;
    for (int64_t pr = blockIdx.x; pr < npairs; pr += gridDim.x, ++np) {
      if (!mbar_wait(BAR_ACC_EMPTY, (uint32_t)((np & 1) ^ 1), s_abort, gerr, 42)) return;
      for (int dh = 0; dh < 4; ++dh, ++n) {
        const int st = n % NSTG;
        if (!mbar_wait(BAR_FULL(st), (uint32_t)((n / NSTG) & 1), s_abort, gerr, 43)) return;
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ads = ad0 + (uint64_t)((st * HALF_BYTES) >> 4);
          const uint64_t bdd = bd0 + (uint64_t)(((dh >> 1) * HID + (dh & 1) * HALF_CHUNKS) * (WT_CHUNK_BYTES >> 4));
#pragma unroll
          for (int j = 0; j < HALF_CHUNKS / 2; ++j)
            mma_tf32_ss(tmem_base, ads + (uint64_t)(j * ((2 * CHUNK_BYTES_A) >> 4)), bdd + (uint64_t)(j * ((2 * WT_CHUNK_BYTES) >> 4)),
                        idesc, (dh | j) ? 1u : 0u);
          mma_commit(BAR_EMPTY(st));
          if (dh == 3) mma_commit(BAR_ACC_FULL);
        }
        __syncwarp();
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int np = 0;
    for (int64_t pr = blockIdx.x; pr < npairs; pr += gridDim.x, ++np) {
      if (!mbar_wait(BAR_ACC_FULL, (uint32_t)(np & 1), s_abort, gerr, 44)) break;
      tc_fence_after();
      float v[48];
#pragma unroll
      for (int c0 = 0; c0 < 48; c0 += 16) {
        float r[16];
        tmem_ld16(taddr + c0, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[c0 + i] = r[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_ACC_EMPTY);
      float4* o4 = reinterpret_cast<float4*>(dx) + (pr * out_chunks + half * KH_CHUNKS) * TM + row;
#pragma unroll
      for (int q = 0; q < KH_CHUNKS; ++q)
        if (half * KH_CHUNKS + q < out_chunks)
          o4[(int64_t)q * TM] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
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
// dW_ih, dW_hh, db of one layer: G[n'][feat] = sum_{t, sample} da[n'] * [x_t | h_prev][feat];  db[n'] = sum da[n'].
//   grid (ctas_per_dir, 2 dirs); each CTA writes one partial block [192 gates][160], reduced by dw_finalize_kernel.
// ---------------------------------------------------------------------------------------------
// K-major formulation (K = the gestures of a tile): both operands must hold FOUR CONSECUTIVE GESTURES of one
// gate / feature row in 16 bytes, i.e. the transpose of the HBM chunk layout (four gates / features of one gesture).
// MN-major TF32 operands (which would read the HBM layout as is) run at about a quarter of the K-major MMA rate
// (DESIGN.md section 4), so sixteen producer warps transpose on the fly instead: coalesced 16-byte loads (lane =
// gesture), a 4x4 transpose inside each lane quad (4 SHFL), conflict-free 16-byte shared stores into the canonical
// no-swizzle K-major layout (k-chunk stride padded by 64 B).  A = da^T (192 gate rows: one M = 128 and one M = 64
// MMA per k-step), B = [x | h_prev | 1 | 0]^T (N = 160: x 0..95, h 96..143, ones row 144 -> bias column).
// Half tiles (64 gestures) double-buffered; accumulators stay in TMEM over all (t, tile) pairs of the persistent CTA.
constexpr int DW_THREADS = 672;                 // warp 0: MMA issuer; warps 1..4: cp.async loaders; warps 5..20: transposers, then read-out
constexpr int DW_LOADERS = 4;
constexpr int DW_N = 160;                       // B rows = columns of the accumulators
constexpr int DW_BIAS_COL = 144;
constexpr int DW_Q = 32;                        // gestures per stage (a quarter tile)
constexpr int DW_KC = DW_Q / 4;                 // k-chunks per stage
constexpr int DW_LBO_A = N4 * 16 + 64;          // 3136 B between k-chunks of the A tile
constexpr int DW_LBO_B = DW_N * 16 + 64;        // 2624 B
constexpr int DW_STAGE = DW_KC * (DW_LBO_A + DW_LBO_B);  // 46080 B
constexpr int DW_RAW_CHUNKS = HID + 24 + KH_CHUNKS;      // 84: da | x | h_prev chunks of one pair
constexpr int DW_RAW_SLOT = DW_RAW_CHUNKS * DW_Q * 16;   // 43008 B: a quarter tile in HBM order
constexpr int DW_NRAW = 3;

__device__ __forceinline__ float4 quad_transpose(float4 v, int lane) {
  // lanes 4q..4q+3 hold rows r0..r3 of a 4x4 block; afterwards lane 4q+j holds column j
  const bool odd = lane & 1, hi = lane & 2;
  float s0 = odd ? v.x : v.y, s1 = odd ? v.z : v.w;
  s0 = __shfl_xor_sync(0xffffffffu, s0, 1);
  s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
  if (odd) { v.x = s0; v.z = s1; } else { v.y = s0; v.w = s1; }
  float t0 = hi ? v.x : v.z, t1 = hi ? v.y : v.w;
  t0 = __shfl_xor_sync(0xffffffffu, t0, 2);
  t1 = __shfl_xor_sync(0xffffffffu, t1, 2);
  if (hi) { v.x = t0; v.y = t1; } else { v.z = t0; v.w = t1; }
  return v;
}

__global__ void __launch_bounds__(DW_THREADS, 1) lstm_tc_dw_kernel(const float* __restrict__ da,
                                                                   const float* __restrict__ xin, int xin_chunks,
                                                                   const float* __restrict__ hself,
                                                                   float* __restrict__ partial, int T, int ntiles,
                                                                   int* __restrict__ gerr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_raw = smem + 2 * DW_STAGE;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_raw + DW_NRAW * DW_RAW_SLOT);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 12);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y;
  const uint32_t bar0 = smem_u32(s_bar);
  auto BAR_FULL = [&](int s) { return bar0 + 8u * s; };            // operand stage filled (16 transposer warps)
  auto BAR_EMPTY = [&](int s) { return bar0 + 16u + 8u * s; };     // operand stage consumed (MMA commit)
  auto BAR_RAW_FULL = [&](int s) { return bar0 + 32u + 8u * s; };  // raw slot landed (bulk-copy bytes)
  auto BAR_RAW_EMPTY = [&](int s) { return bar0 + 56u + 8u * s; }; // raw slot read by all transposer warps
  const uint32_t BAR_DONE = bar0 + 80u;
  {
    float4* z = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < 2 * DW_STAGE / 16; i += DW_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    // the ones row of B (bias column of the accumulators): never overwritten by the transposers
    for (int i = tid; i < 2 * DW_KC; i += DW_THREADS) {
      const int st = i / DW_KC, kc = i % DW_KC;
      *reinterpret_cast<float4*>(smem + st * DW_STAGE + DW_KC * DW_LBO_A + kc * DW_LBO_B + (DW_BIAS_COL / 8) * 128 +
                                 (DW_BIAS_COL % 8) * 16) = make_float4(1.f, 1.f, 1.f, 1.f);
    }
  }
  if (tid == 0) {
    for (int st = 0; st < 2; ++st) { mbar_init(BAR_FULL(st), 16); mbar_init(BAR_EMPTY(st), 1); }
    for (int st = 0; st < DW_NRAW; ++st) { mbar_init(BAR_RAW_FULL(st), DW_LOADERS * 32); mbar_init(BAR_RAW_EMPTY(st), 16); }
    mbar_init(BAR_DONE, 1);
    *s_abort = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int64_t npairs = (int64_t)T * ntiles;
  const int nchunks = HID + xin_chunks + KH_CHUNKS;

  if (warp == 0) {
    // MMA issuer: warp-uniform loop, one elected lane issues
    const uint32_t id128 = make_idesc(128, DW_N), id64 = make_idesc(64, DW_N);
    int n = 0;
    bool ok = true;
    for (int64_t pr = blockIdx.x; pr < npairs && ok; pr += gridDim.x) {
      for (int q = 0; q < TM / DW_Q; ++q, ++n) {
        const int st = n & 1;
        if (!mbar_wait(BAR_FULL(st), (uint32_t)((n >> 1) & 1), s_abort, gerr, 51)) { ok = false; break; }
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem) + st * DW_STAGE, b0 = a0 + DW_KC * DW_LBO_A;
        const uint64_t ad0 = make_desc(a0, DW_LBO_A, 128), ad1 = make_desc(a0 + 16 * 128, DW_LBO_A, 128);
        const uint64_t bd0 = make_desc(b0, DW_LBO_B, 128);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < DW_Q / 8; ++ks) {
            const uint32_t acc = (n | ks) ? 1u : 0u;
            const uint64_t astep = (uint64_t)(ks * ((2 * DW_LBO_A) >> 4)), bstep = (uint64_t)(ks * ((2 * DW_LBO_B) >> 4));
            mma_tf32_ss(tmem_base, ad0 + astep, bd0 + bstep, id128, acc);
            mma_tf32_ss(tmem_base + (uint32_t)DW_N, ad1 + astep, bd0 + bstep, id64, acc);
          }
          mma_commit(BAR_EMPTY(st));
        }
        __syncwarp();
      }
    }
    if (ok && elect_one()) mma_commit(BAR_DONE);
  } else if (warp <= DW_LOADERS) {
    // loaders: 16-byte cp.async copies in HBM order (one warp instruction = the 32 gestures of one chunk), completion
    // signalled on the slot's mbarrier by every thread; never blocks on data, so up to three quarter tiles are in
    // flight.  (512-byte bulk copies were issue-bound here: ~57 clk each from a single warp.)
    const int lw = warp - 1;
    int n = 0;
    bool ok = true;
    for (int64_t pr = blockIdx.x; pr < npairs && ok; pr += gridDim.x) {
      const int t = (int)(pr / ntiles), tile = (int)(pr % ntiles);
      const int tp = dir ? t + 1 : t - 1;  // timestep whose h fed the recurrence at t
      const bool has_prev = tp >= 0 && tp < T;
      const float4* sd = reinterpret_cast<const float4*>(da) + ((int64_t)dir * npairs + pr) * HID * TM + lane;
      const float4* sx = reinterpret_cast<const float4*>(xin) + pr * xin_chunks * TM + lane;
      const float4* sh = reinterpret_cast<const float4*>(hself) +
                         (((int64_t)(has_prev ? tp : 0) * ntiles + tile) * (2 * KH_CHUNKS) + dir * KH_CHUNKS) * TM + lane;
      const int nload = has_prev ? nchunks : nchunks - KH_CHUNKS;
      for (int q = 0; q < TM / DW_Q; ++q, ++n) {
        const int slot = n % DW_NRAW;
        if (!mbar_wait(BAR_RAW_EMPTY(slot), (uint32_t)(((n / DW_NRAW) & 1) ^ 1), s_abort, gerr, 54)) { ok = false; break; }
        const uint32_t dst0 = smem_u32(s_raw) + slot * DW_RAW_SLOT + lane * 16;
        for (int c = lw; c < nload; c += DW_LOADERS) {
          const float4* src = c < HID ? sd + c * TM : (c < HID + xin_chunks ? sx + (c - HID) * TM : sh + (c - HID - xin_chunks) * TM);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + c * (DW_Q * 16)), "l"(src + q * DW_Q) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];" ::"r"(BAR_RAW_FULL(slot)) : "memory");
      }
    }
  } else {
    const int tw = warp - 1 - DW_LOADERS;  // 0..15
    int n = 0;
    bool ok = true;
    for (int64_t pr = blockIdx.x; pr < npairs && ok; pr += gridDim.x) {
      const int t = (int)(pr / ntiles);
      const int tp = dir ? t + 1 : t - 1;
      const bool has_prev = tp >= 0 && tp < T;
      for (int q = 0; q < TM / DW_Q; ++q, ++n) {
        const int st = n & 1, slot = n % DW_NRAW;
        if (!mbar_wait(BAR_RAW_FULL(slot), (uint32_t)((n / DW_NRAW) & 1), s_abort, gerr, 55)) { ok = false; break; }
        if (!mbar_wait(BAR_EMPTY(st), (uint32_t)(((n >> 1) & 1) ^ 1), s_abort, gerr, 52)) { ok = false; break; }
        uint8_t* sa = smem + st * DW_STAGE;
        uint8_t* sb = sa + DW_KC * DW_LBO_A;
        const float4* raw = reinterpret_cast<const float4*>(s_raw + slot * DW_RAW_SLOT) + lane;
        const int kc = lane >> 2, j = lane & 3;
#pragma unroll 2
        for (int c = tw; c < nchunks; c += 16) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < HID + xin_chunks || has_prev) v = raw[c * DW_Q];
          const float4 w = quad_transpose(v, lane);
          int row;
          uint8_t* dst;
          if (c < HID) { row = 4 * c + j; dst = sa + kc * DW_LBO_A; }
          else if (c < HID + xin_chunks) { row = 4 * (c - HID) + j; dst = sb + kc * DW_LBO_B; }
          else { row = 96 + 4 * (c - HID - xin_chunks) + j; dst = sb + kc * DW_LBO_B; }
          *reinterpret_cast<float4*>(dst + (row >> 3) * 128 + (row & 7) * 16) = w;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(BAR_FULL(st));
          mbar_arrive(BAR_RAW_EMPTY(slot));
        }
      }
    }
    if (ok && mbar_wait(BAR_DONE, 0, s_abort, gerr, 53)) {
      tc_fence_after();
      // accumulators: gates 0..127 in TMEM lanes 0..127 (cols 0..159); gates 128..191 as an M = 64 tile (lanes 0..15
      // of each quarter, cols 160..319).  Four warps per quarter split the 160 columns.
      const int quarter = warp & 3, cpart = tw >> 2;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      float* pbase = partial + ((int64_t)dir * gridDim.x + blockIdx.x) * N4 * DW_N;
      for (int c0 = cpart * 40; c0 < cpart * 40 + 40; c0 += 8) {
        float r[8];
        tmem_ld8(taddr + c0, r);
        float* d0 = pbase + (int64_t)(quarter * 32 + lane) * DW_N + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i) d0[i] = r[i];
        tmem_ld8(taddr + DW_N + c0, r);
        if (lane < 16) {
          float* d1 = pbase + (int64_t)(128 + quarter * 16 + lane) * DW_N + c0;
#pragma unroll
          for (int i = 0; i < 8; ++i) d1[i] = r[i];
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

// reduce the per-CTA partials into the flat gradient (+=), mapping n' = 4u + g back to PyTorch rows g*H + u
__global__ void dw_finalize_kernel(const float* __restrict__ partial, int nparts, int I, float* __restrict__ dlp,
                                   int64_t dir_stride, int64_t off_whh, int64_t off_bih, int64_t off_bhh) {
  const int dir = blockIdx.y;
  const int per_row = I + HID + 1;
  float* o = dlp + dir * dir_stride;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N4 * per_row; idx += gridDim.x * blockDim.x) {
    const int np = idx / per_row, c = idx % per_row;
    const int u = np >> 2, g = np & 3, prow = g * HID + u;
    const int col = c < I ? c : (c < I + HID ? 96 + (c - I) : DW_BIAS_COL);
    float s = 0.f;
#pragma unroll 8
    for (int p = 0; p < nparts; ++p) s += __ldg(partial + (((int64_t)dir * nparts + p) * N4 + np) * DW_N + col);
    if (c < I) o[(int64_t)prow * I + c] += s;
    else if (c < I + HID) o[off_whh + (int64_t)prow * HID + (c - I)] += s;
    else { o[off_bih + prow] += s; o[off_bhh + prow] += s; }
  }
}

// [T][B][W] row-major  ->  [T][ntiles][W/4][128][4] chunk layout (rows beyond B are zero)
__global__ void rows_to_chunk_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int64_t B,
                                     int ntiles, int W) {
  const int64_t n = (int64_t)T * ntiles * TM * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i & 3);
    const int row = (int)((i >> 2) % TM);
    const int q = (int)((i / (4 * TM)) % (W / 4));
    const int64_t pr = i / ((int64_t)W * TM);
    const int tile = (int)(pr % ntiles), t = (int)(pr / ntiles);
    const int64_t b = (int64_t)tile * TM + row;
    out[i] = b < B ? __ldg(in + ((int64_t)t * B + b) * W + q * 4 + c4) : 0.f;
  }
}

// Output head backward, fused (models.py:160-163): per (timestep, 128-gesture tile)
//   dpre = dy * (1 - y^2);   dh = dpre * Wo  -> chunk layout (the top LSTM layer's dh);
//   dWo += dpre^T * h,  dbo += sum dpre      -> per-CTA partials, reduced in fixed order by head_bwd_finalize_kernel.
// Phase A: thread = gesture row (dpre to smem, 24 coalesced 16-byte stores of dh).  Phase B: thread = feature
// (reads the top layer's output in its chunk layout - the same TF32-rounded h the forward head multiplied - three
// running sums).  HBM-bound: h read once, dh written once.
constexpr int HEAD_MAXC = 4;
__global__ void __launch_bounds__(128) head_bwd_tc_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                          const float* __restrict__ h_tc, const float* __restrict__ wo,
                                                          float* __restrict__ dh, float* __restrict__ part, int T,
                                                          int64_t B, int ntiles, int C) {
  __shared__ float s_wo[HEAD_MAXC * 96];
  __shared__ float s_dp[TM * HEAD_MAXC];
  const int tid = threadIdx.x;
  for (int i = tid; i < HEAD_MAXC * 96; i += 128) s_wo[i] = i < C * 96 ? __ldg(wo + i) : 0.f;
  float accw[HEAD_MAXC] = {0.f, 0.f, 0.f, 0.f};
  float accb = 0.f;  // thread c < C sums dpre[:, c]
  const int64_t npairs = (int64_t)T * ntiles;
  __syncthreads();
  for (int64_t pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
    const int t = (int)(pr / ntiles), tile = (int)(pr % ntiles);
    const int64_t b = (int64_t)tile * TM + tid;
    float dp[HEAD_MAXC] = {0.f, 0.f, 0.f, 0.f};
    if (b < B) {
      const int64_t src = (b * T + t) * C;
      for (int c = 0; c < C; ++c) {
        const float yy = __ldg(y + src + c);
        dp[c] = __ldg(dy + src + c) * (1.f - yy * yy);
      }
    }
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c) s_dp[tid * HEAD_MAXC + c] = dp[c];
    float4* o4 = reinterpret_cast<float4*>(dh) + pr * 24 * TM + tid;
#pragma unroll 4
    for (int q = 0; q < 24; ++q) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < HEAD_MAXC; ++c) a = fmaf(dp[c], s_wo[c * 96 + q * 4 + i], a);
        v[i] = a;
      }
      o4[(int64_t)q * TM] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    const int64_t b0 = (int64_t)tile * TM;
    const int nrow = (int)((B - b0) < TM ? (B - b0) : TM);
    if (tid < 96) {
      // feature tid of gesture r: chunk tid / 4, component tid % 4 of the chunk-layout tile [24][128][4]
      const float* hp = h_tc + (pr * 24 + (tid >> 2)) * (int64_t)(TM * 4) + (tid & 3);
#pragma unroll 8
      for (int r = 0; r < nrow; ++r) {
        const float hv = __ldg(hp + r * 4);
        const float4 d = *reinterpret_cast<const float4*>(s_dp + r * HEAD_MAXC);
        accw[0] = fmaf(d.x, hv, accw[0]);
        accw[1] = fmaf(d.y, hv, accw[1]);
        accw[2] = fmaf(d.z, hv, accw[2]);
        accw[3] = fmaf(d.w, hv, accw[3]);
      }
    } else if (tid - 96 < C) {
      for (int r = 0; r < nrow; ++r) accb += s_dp[r * HEAD_MAXC + (tid - 96)];
    }
    __syncthreads();
  }
  // partial layout per CTA: [C][96] weights then [C] bias
  float* pp = part + (int64_t)blockIdx.x * (HEAD_MAXC * 96 + HEAD_MAXC);
  if (tid < 96) {
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c) pp[c * 96 + tid] = accw[c];
  } else if (tid - 96 < HEAD_MAXC) {
    pp[HEAD_MAXC * 96 + (tid - 96)] = accb;
  }
}

// one block per output element (C*96 weights + C biases): fixed-order tree sum over the CTA partials, accumulated
// into the gradient buffer
__global__ void __launch_bounds__(128) head_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int C,
                                                                float* __restrict__ dwo, float* __restrict__ dbo) {
  __shared__ float s[128];
  const int o = blockIdx.x;  // < C*96: weight (c = o / 96, f = o % 96); else bias c = o - C*96
  const int idx = o < C * 96 ? o : HEAD_MAXC * 96 + (o - C * 96);
  float a = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 128) a += part[(int64_t)i * (HEAD_MAXC * 96 + HEAD_MAXC) + idx];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int w = 64; w > 0; w >>= 1) {
    if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (o < C * 96) dwo[o] += s[0];
    else dbo[o - C * 96] += s[0];
  }
}

// dz[b][j] = sum_t dx0[t][tile][chunk][row][.] at feature pd + j     (backward of repeat + cat, models.py:154-157)
__global__ void dz_chunk_kernel(const float* __restrict__ dx0, float* __restrict__ dz, int T, int64_t B, int ntiles,
                                int chunks, int pd, int Z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Z) return;
  const int64_t b = i / Z;
  const int f = pd + (int)(i % Z);
  const int tile = (int)(b / TM), row = (int)(b % TM);
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += __ldg(dx0 + ((((int64_t)t * ntiles + tile) * chunks + (f >> 2)) * TM + row) * 4 + (f & 3));
  dz[i] = s;
}

// out[b][t][:] = tanh(W_o [h_fwd, h_bwd] + b_o) from the last layer's tc-layout output   (models.py:163)
// block: 8 warps = 8 timesteps x 32 rows; results staged in smem so each row writes 8*C contiguous floats.
__global__ void __launch_bounds__(256) head_tc_kernel(const float* __restrict__ h, const float* __restrict__ wo,
                                                      const float* __restrict__ bo, float* __restrict__ out, int T,
                                                      int64_t B, int ntiles, int C) {
  __shared__ float s_w[3 * 96 + 3];
  __shared__ float s_o[32][8 * 3 + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < C * 96 + C; i += 256) s_w[i] = i < C * 96 ? wo[i] : bo[i - C * 96];
  __syncthreads();
  const int64_t row0 = (int64_t)blockIdx.x * 32;
  const int t0 = blockIdx.y * 8;
  const int t = t0 + warp;
  const int64_t b = row0 + lane;
  const int tile = (int)(b / TM), r = (int)(b % TM);
  if (t < T && tile < ntiles) {
    const float4* hp = reinterpret_cast<const float4*>(h) + ((int64_t)t * ntiles + tile) * 24 * TM + r;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll 4
    for (int ck = 0; ck < 24; ++ck) {
      const float4 v = __ldg(hp + (int64_t)ck * TM);
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const float* w = s_w + cc * 96 + ck * 4;
        acc[cc] = fmaf(v.x, w[0], fmaf(v.y, w[1], fmaf(v.z, w[2], fmaf(v.w, w[3], acc[cc]))));
      }
    }
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) s_o[lane][warp * 3 + cc] = tanhf(acc[cc] + s_w[C * 96 + cc]);
  }
  __syncthreads();
  // 32 rows x (8 timesteps x 3) floats; each row's 24 floats are contiguous in out
  for (int i = tid; i < 32 * 24; i += 256) {
    const int rr = i / 24, j = i % 24;
    const int64_t bb = row0 + rr;
    const int tt = t0 + j / 3;
    if (bb < B && tt < T) out[(bb * T + tt) * C + (j % 3)] = s_o[rr][j];
  }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kKX0 = 40;  // layer-0 K (2 + 32 = 34, or 35 with time) padded to a multiple of 8

struct TcPlan {
  int T, L, Z, pd, C;
  int ntiles;
  int64_t rows;
  int64_t x0_floats, h_floats, img_floats[WGG_MAX_HIDDEN_LAYERS], img_off[WGG_MAX_HIDDEN_LAYERS], total;
};

bool tc_plan(const wgg_model_cfg* c, int64_t B, TcPlan* p) {
  if (c->gen_hidden_dim != tc::HID || c->input_dim != 3) return false;
  const int pd = c->prototype_has_time ? 3 : 2;
  if (pd + c->latent_dim > kKX0) return false;
  p->T = c->seq_length; p->L = c->gen_num_layers; p->Z = c->latent_dim; p->pd = pd; p->C = c->input_dim;
  p->ntiles = (int)cdiv64(B, tc::TM);
  p->rows = (int64_t)p->ntiles * tc::TM;
  p->x0_floats = (int64_t)p->T * p->rows * kKX0;
  p->h_floats = (int64_t)p->T * p->rows * 96;
  int64_t off = p->x0_floats + 2 * p->h_floats;
  for (int l = 0; l < p->L; ++l) {
    const int KX = l == 0 ? kKX0 : 96;
    p->img_floats[l] = (int64_t)tc::N4 * (KX + tc::HID) + tc::N4;
    p->img_off[l] = off;
    off += 2 * p->img_floats[l];
  }
  p->total = off;
  return true;
}

template <int KXC, int STASH, int NCH = 3>
int launch_layer(wgg_ctx* ctx, const float* xin, const float* img, int64_t img_stride, float* hout, int T, int ntiles,
                 int64_t B, float* gc, float* h_rm, cudaStream_t st) {
  constexpr size_t smem = (size_t)KXC * tc::CHUNK_BYTES_W + tc::KH_CHUNKS * tc::CHUNK_BYTES_W + tc::NSTAGE * tc::SB_BYTES +
                          tc::KH_CHUNKS * tc::CHUNK_BYTES_A + tc::N4 * 4 + 20 * 8 + 16;
  if (!wgg_smem_ok(ctx, tc::lstm_tc_fwd_kernel<KXC, STASH, NCH>, smem))
    return wgg_fail(ctx, WGG_ECUDA, "lstm_tc_fwd_kernel: cannot reserve shared memory%s");
  dim3 grid((unsigned)ntiles, 2);
  // algorithmic FLOPs: 2 dirs x T x B x 2 x 192 x (K_x + 48); bytes: x in (both dirs read it) + h out, and for the
  // grad-carrying forward the stash it has to write (gates + c: 960 B per gesture-step-direction, + h_rm rows)
  ProfScope prof(ctx, "lstm_tc_fwd_kernel", st, 2.0 * T * (double)B * 2.0 * tc::N4 * (KXC * 4 + tc::HID),
                 (double)T * B * 4.0 * (2.0 * KXC * 4 + 96) +
                     (STASH ? (double)T * B * (2.0 * tc::GC_CHUNKS * 16 + (h_rm ? 96 * 4.0 : 0.0)) : 0.0));
  tc::lstm_tc_fwd_kernel<KXC, STASH, NCH><<<grid, tc::fwd_threads(NCH), smem, st>>>(xin, img, img_stride, hout, T, ntiles, gc, h_rm, B,
                                                                         ctx->async_err);
  WGG_CHECK_LAUNCH(ctx, "lstm_tc_fwd_kernel");
  return WGG_OK;
}
}  // namespace

// ---- training stash of the tcgen05 path (floats):  x0_tc | h_tc[0..L-1] | gc[0..L-1] ----
struct TcStash {
  int64_t x0, h[WGG_MAX_HIDDEN_LAYERS], gc[WGG_MAX_HIDDEN_LAYERS], total;
};
static void tc_stash_layout(const TcPlan& p, int64_t B, TcStash* s) {
  int64_t off = 0;
  s->x0 = off; off += p.x0_floats;
  for (int l = 0; l < p.L; ++l) { s->h[l] = off; off += p.h_floats; }
  for (int l = 0; l < p.L; ++l) { s->gc[l] = off; off += (int64_t)2 * p.T * p.rows * (tc::GC_CHUNKS * 4); }
  s->total = off;
}

int64_t generator_tc_stash_floats(const wgg_model_cfg* cfg, int64_t B) {
  TcPlan p;
  if (!tc_plan(cfg, B, &p)) return 0;
  TcStash s;
  tc_stash_layout(p, B, &s);
  return s.total;
}

int64_t generator_tc_workspace_floats(const wgg_model_cfg* cfg, int64_t B) {
  TcPlan p;
  return tc_plan(cfg, B, &p) ? p.total : 0;
}

// backward workspace: dh ping-pong (2 x T*R*96) | da (2*T*R*192) | W images | dW partials
static const int kDwCtasPerDir = 74;
int64_t generator_tc_bwd_workspace_floats(const wgg_model_cfg* cfg, int64_t B) {
  TcPlan p;
  if (!tc_plan(cfg, B, &p)) return 0;
  return 2 * p.h_floats + (int64_t)2 * p.T * p.rows * tc::N4 + 2 * tc::HID * tc::N4 + 4 * tc::HID * tc::N4 +
         (int64_t)2 * kDwCtasPerDir * tc::N4 * tc::DW_N;
}

bool generator_tc_supported(const wgg_model_cfg* cfg) {
  TcPlan p;
  return tc_plan(cfg, 1, &p);
}

// Generator forward on the tcgen05 path.  stash == nullptr: no-grad (activations ping-pong in ws); otherwise the
// grad-carrying forward keeps every layer's input/output (chunk layout), gates and cell states for BPTT.
int generator_forward_tc(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const int64_t* layer_off,
                         const int64_t* dir_stride, const int64_t* off_whh, const int64_t* off_bih,
                         const int64_t* off_bhh, int64_t off_wo, int64_t off_bo, const float* proto, const float* z,
                         int64_t B, float* out, float* ws, int64_t ws_floats, float* stash, cudaStream_t st) {
  TcPlan p;
  if (!tc_plan(cfg, B, &p)) return wgg_fail(ctx, WGG_EUNSUPPORTED, "generator_forward_tc: unsupported configuration%s");
  if (!ws || ws_floats < p.total) return wgg_fail(ctx, WGG_EWORKSPACE, "generator_forward_tc: workspace too small%s");
  TcStash sl;
  tc_stash_layout(p, B, &sl);
  float* x0 = stash ? stash + sl.x0 : ws;
  float* hbuf[2] = {ws + p.x0_floats, ws + p.x0_floats + p.h_floats};
  for (int l = 0; l < p.L; ++l) {
    const int I = l == 0 ? p.pd + p.Z : 96;
    const int KX = l == 0 ? kKX0 : 96;
    tc::prep_weights_kernel<<<dim3(32, 2), 256, 0, st>>>(params + layer_off[l], dir_stride[l], off_whh[l], off_bih[l],
                                                         off_bhh[l], I, KX, ws + p.img_off[l], p.img_floats[l]);
    WGG_CHECK_LAUNCH(ctx, "prep_weights_kernel");
  }
  tc::build_x0_tc_kernel<<<dim3((unsigned)p.ntiles, (unsigned)((p.T + 7) / 8)), 256, (size_t)tc::TM * (p.Z + 1) * sizeof(float), st>>>(
      proto, z, x0, p.T, B, p.ntiles, p.C, p.pd, p.Z, kKX0);
  WGG_CHECK_LAUNCH(ctx, "build_x0_tc_kernel");
  const float* in = x0;
  for (int l = 0; l < p.L; ++l) {
    float* hout = stash ? stash + sl.h[l] : hbuf[l & 1];
    float* gcl = stash ? stash + sl.gc[l] : nullptr;
    float* hrm = nullptr;  // the row-major copy of the top layer's output is no longer needed (head backward reads the chunk layout)
    const float* img = ws + p.img_off[l];
    if (l == 0) {
      if (stash) WGG_TRY((launch_layer<kKX0 / 4, 1>(ctx, in, img, p.img_floats[l], hout, p.T, p.ntiles, B, gcl, hrm, st)));
      else WGG_TRY((launch_layer<kKX0 / 4, 0, 3>(ctx, in, img, p.img_floats[l], hout, p.T, p.ntiles, B, nullptr, nullptr, st)));
    } else {
      if (stash) WGG_TRY((launch_layer<24, 1>(ctx, in, img, p.img_floats[l], hout, p.T, p.ntiles, B, gcl, hrm, st)));
      else WGG_TRY((launch_layer<24, 0, 3>(ctx, in, img, p.img_floats[l], hout, p.T, p.ntiles, B, nullptr, nullptr, st)));
    }
    in = hout;
  }
  dim3 grid((unsigned)cdiv64(B, 32), (unsigned)((p.T + 7) / 8));
  tc::head_tc_kernel<<<grid, 256, 0, st>>>(in, params + off_wo, params + off_bo, out, p.T, B, p.ntiles, p.C);
  WGG_CHECK_LAUNCH(ctx, "head_tc_kernel");
  return WGG_OK;
}

// LSTM stack backward on the tcgen05 path.  dh_rm = d(last layer output) [T][B][2H] row-major (from the head's
// backward); accumulates all LSTM parameter gradients into dparams and writes dz (may be null).
int generator_backward_tc_layers(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, float* dparams,
                                 const int64_t* layer_off, const int64_t* dir_stride, const int64_t* off_whh,
                                 const int64_t* off_bih, const int64_t* off_bhh, int64_t B, const float* stash,
                                 const float* out, const float* dout, int64_t off_wo, int64_t off_bo, float* dz,
                                 float* ws, int64_t ws_floats, cudaStream_t st) {
  TcPlan p;
  if (!tc_plan(cfg, B, &p)) return wgg_fail(ctx, WGG_EUNSUPPORTED, "generator_backward_tc: unsupported configuration%s");
  if (!ws || ws_floats < generator_tc_bwd_workspace_floats(cfg, B))
    return wgg_fail(ctx, WGG_EWORKSPACE, "generator_backward_tc: workspace too small%s");
  TcStash sl;
  tc_stash_layout(p, B, &sl);
  float* dh[2] = {ws, ws + p.h_floats};
  float* da = dh[1] + p.h_floats;
  float* whhT = da + (int64_t)2 * p.T * p.rows * tc::N4;
  float* wihT = whhT + 2 * tc::HID * tc::N4;
  float* part = wihT + 4 * tc::HID * tc::N4;
  {
    // output head backward: dh of the top layer straight into the chunk layout, dWo/dbo through per-CTA partials
    if (cfg->input_dim > tc::HEAD_MAXC) return wgg_fail(ctx, WGG_EUNSUPPORTED, "generator_backward_tc: input_dim > 4%s");
    const int64_t npairs0 = (int64_t)p.T * p.ntiles;
    const int hgrid = (int)(npairs0 < 8 * (int64_t)ctx->sm_count ? npairs0 : 8 * (int64_t)ctx->sm_count);
    ProfScope prof(ctx, "head_bwd_tc_kernel", st, 4.0 * p.T * (double)B * 96 * cfg->input_dim, p.T * (double)B * 4.0 * (96 + 96));
    tc::head_bwd_tc_kernel<<<hgrid, 128, 0, st>>>(out, dout, stash + sl.h[p.L - 1], params + off_wo, dh[0], part, p.T, B, p.ntiles,
                                                  cfg->input_dim);
    WGG_CHECK_LAUNCH(ctx, "head_bwd_tc_kernel");
    tc::head_bwd_finalize_kernel<<<cfg->input_dim * 96 + cfg->input_dim, 128, 0, st>>>(part, hgrid, cfg->input_dim,
                                                                                      dparams + off_wo, dparams + off_bo);
    WGG_CHECK_LAUNCH(ctx, "head_bwd_finalize_kernel");
  }
  int cur = 0;
  constexpr size_t smem_bwd = (size_t)tc::HID * tc::CHUNK_BYTES_A + tc::HID * tc::WT_CHUNK_BYTES + 4 * 8 + 16;
  constexpr size_t smem_dx = (size_t)3 * (tc::HID / 2) * tc::CHUNK_BYTES_A + 2 * tc::HID * tc::WT_CHUNK_BYTES + 8 * 8 + 16;
  constexpr size_t smem_dw = (size_t)2 * tc::DW_STAGE + tc::DW_NRAW * tc::DW_RAW_SLOT + 12 * 8 + 16;
  if (!wgg_smem_ok(ctx, tc::lstm_tc_bwd_kernel, smem_bwd) || !wgg_smem_ok(ctx, tc::lstm_tc_dx_kernel, smem_dx) ||
      !wgg_smem_ok(ctx, tc::lstm_tc_dw_kernel, smem_dw))
    return wgg_fail(ctx, WGG_ECUDA, "generator_backward_tc: cannot reserve shared memory%s");
  const int64_t npairs = (int64_t)p.T * p.ntiles;
  for (int l = p.L - 1; l >= 0; --l) {
    const int I = l == 0 ? p.pd + p.Z : 96;
    const float* lp = params + layer_off[l];
    float* dlp = dparams + layer_off[l];
    tc::prep_whhT_kernel<<<dim3(8, 2), 256, 0, st>>>(lp, dir_stride[l], off_whh[l], whhT);
    WGG_CHECK_LAUNCH(ctx, "prep_whhT_kernel");
    {
      dim3 grid((unsigned)p.ntiles, 2);
      ProfScope prof(ctx, "lstm_tc_bwd_kernel", st, 2.0 * p.T * (double)B * 2.0 * tc::N4 * tc::HID,
                     2.0 * p.T * (double)B * 4.0 * (tc::GC_CHUNKS * 4 + tc::HID + tc::HID + tc::N4));
      tc::lstm_tc_bwd_kernel<<<grid, tc::BWD_THREADS, smem_bwd, st>>>(stash + sl.gc[l], whhT, dh[cur], da, p.T, p.ntiles, B,
                                                                      ctx->async_err);
      WGG_CHECK_LAUNCH(ctx, "lstm_tc_bwd_kernel");
    }
    {
      const float* xin = l == 0 ? stash + sl.x0 : stash + sl.h[l - 1];
      const int xin_chunks = l == 0 ? kKX0 / 4 : 24;
      dim3 grid((unsigned)(npairs < kDwCtasPerDir ? npairs : kDwCtasPerDir), 2);
      ProfScope prof(ctx, "lstm_tc_dw_kernel", st, 2.0 * p.T * (double)B * 2.0 * tc::N4 * (I + tc::HID),
                     2.0 * p.T * (double)B * 4.0 * (tc::N4 + I + tc::HID));
      tc::lstm_tc_dw_kernel<<<grid, tc::DW_THREADS, smem_dw, st>>>(da, xin, xin_chunks, stash + sl.h[l], part, p.T, p.ntiles,
                                                                   ctx->async_err);
      WGG_CHECK_LAUNCH(ctx, "lstm_tc_dw_kernel");
      tc::dw_finalize_kernel<<<dim3(72, 2), 256, 0, st>>>(part, (int)grid.x, I, dlp, dir_stride[l], off_whh[l], off_bih[l],
                                                          off_bhh[l]);
      WGG_CHECK_LAUNCH(ctx, "dw_finalize_kernel");
    }
    if (l > 0 || dz) {
      const int halves = l == 0 ? 1 : 2;
      const int out_chunks = l == 0 ? kKX0 / 4 : 24;
      tc::prep_wihT_kernel<<<dim3(8, 2, halves), 256, 0, st>>>(lp, dir_stride[l], I, wihT);
      WGG_CHECK_LAUNCH(ctx, "prep_wihT_kernel");
      int gx = ctx->sm_count / halves;
      if (gx > npairs) gx = (int)npairs;
      dim3 grid((unsigned)gx, (unsigned)halves);
      ProfScope prof(ctx, "lstm_tc_dx_kernel", st, 2.0 * p.T * (double)B * 2.0 * tc::N4 * I,
                     p.T * (double)B * 4.0 * (2.0 * tc::N4 + I));
      tc::lstm_tc_dx_kernel<<<grid, tc::DX_THREADS, smem_dx, st>>>(da, wihT, dh[cur ^ 1], p.T, p.ntiles, out_chunks,
                                                                   ctx->async_err);
      WGG_CHECK_LAUNCH(ctx, "lstm_tc_dx_kernel");
      cur ^= 1;
    }
  }
  if (dz) {
    tc::dz_chunk_kernel<<<(unsigned)cdiv64(B * p.Z, 256), 256, 0, st>>>(dh[cur], dz, p.T, B, p.ntiles, kKX0 / 4, p.pd, p.Z);
    WGG_CHECK_LAUNCH(ctx, "dz_chunk_kernel");
  }
  return WGG_OK;
}

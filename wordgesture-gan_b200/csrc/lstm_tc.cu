// Persistent BiLSTM-layer forward on 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces, for the no-grad generator calls (10 of the 12 per training batch: src/shared/utils.py:70-73,91-94;
// and all of sampling: eval_gan.py:131-135), nn.LSTM's per-layer work of src/gan/models.py:160:
//     gates_t = x_t W_ih^T + h_{t-1} W_hh^T + b ;  c_t = f c_{t-1} + i g ;  h_t = o tanh(c_t)
// One CTA owns a tile of 128 samples of one direction for all T timesteps:
//   * W_ih, W_hh (TF32) stay resident in shared memory for the whole sequence;
//   * x_t tiles stream in through a 5-stage ring of bulk asynchronous copies (cp.async.bulk -> mbarrier tx);
//   * one elected thread issues tcgen05.mma kind::tf32 (M=128 samples, N=192 gates, K=8 per instruction), the
//     accumulator lives in TMEM (two 192-column buffers: the x-projection of step t+1 is issued while the
//     epilogue of step t still runs; only h_{t-1} W_hh^T sits on the recurrent critical path);
//   * 8 epilogue warps read the accumulator with tcgen05.ld (one thread = one sample row, gates of a hidden
//     unit are adjacent columns because the weight rows are permuted unit-major), apply the gate
//     non-linearities, keep c in registers, and write h_t both to shared memory (the next step's A operand)
//     and to HBM (the next layer's input).
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
constexpr int NTHREADS = 320;                  // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
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
__global__ void build_x0_tc_kernel(const float* __restrict__ proto, const float* __restrict__ z, float* __restrict__ x0,
                                   int T, int64_t B, int ntiles, int C, int pd, int Z, int KX0) {
  const int64_t per_t = (int64_t)ntiles * TM * KX0;
  const int64_t n = (int64_t)T * per_t;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // enumerate in storage order so that writes are coalesced
    const int t = (int)(i / per_t);
    int64_t r = i % per_t;
    const int tile = (int)(r / ((int64_t)TM * KX0));
    r %= (int64_t)TM * KX0;
    const int chunk = (int)(r / (TM * 4));
    const int row = (int)((r % (TM * 4)) / 4);
    const int k = chunk * 4 + (int)(r & 3);
    const int64_t b = (int64_t)tile * TM + row;
    float v = 0.f;
    if (b < B) {
      if (k < pd) v = __ldg(proto + (b * T + t) * C + k);
      else if (k < pd + Z) v = __ldg(z + b * Z + (k - pd));
    }
    x0[i] = rna_tf32(v);
  }
}

// ---------------------------------------------------------------------------------------------
// the persistent layer kernel.  grid (ntiles, 2 directions), 320 threads, 1 CTA / SM.
// xin : [T][ntiles][KXC][16][8][4]   hout : [T][ntiles][24][16][8][4] (this direction fills chunks dir*12..+12)
// ---------------------------------------------------------------------------------------------
// STASH = 1 (grad-carrying forward) additionally writes, per step, what BPTT needs:
//   gc   : [2 dirs][T][ntiles][GC_CHUNKS][128 rows][4]  chunk u < 48 = (i,f,g,o) of hidden unit u, chunk 48 + u/4 = c
//   h_rm : [T][B][2H] row-major, un-rounded fp32 (operand of the weight-gradient GEMMs and of the output head)
constexpr int GC_CHUNKS = HID + HID / 4;  // 60

template <int KXC, int STASH>
__global__ void __launch_bounds__(NTHREADS, 1) lstm_tc_fwd_kernel(const float* __restrict__ xin,
                                                                  const float* __restrict__ wimg, int64_t img_stride,
                                                                  float* __restrict__ hout, int T, int ntiles,
                                                                  float* __restrict__ gc, float* __restrict__ h_rm,
                                                                  int64_t B, int* __restrict__ gerr) {
  constexpr int NSB = (KXC + SB_CHUNKS - 1) / SB_CHUNKS;
  constexpr int WX_BYTES = KXC * CHUNK_BYTES_W;
  constexpr int WH_BYTES = KH_CHUNKS * CHUNK_BYTES_W;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_wx = smem;
  uint8_t* s_wh = s_wx + WX_BYTES;
  uint8_t* s_x = s_wh + WH_BYTES;
  uint8_t* s_h = s_x + NSTAGE * SB_BYTES;
  float* s_bias = reinterpret_cast<float*>(s_h + KH_CHUNKS * CHUNK_BYTES_A);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + N4);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 16);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, dir = blockIdx.y;
  const uint32_t bar0 = smem_u32(s_bar);
  auto BAR_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto BAR_ACC_FULL = [&](int b) { return bar0 + 8u * (2 * NSTAGE + b); };
  auto BAR_ACC_EMPTY = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 2 + b); };
  const uint32_t BAR_H = bar0 + 8u * (2 * NSTAGE + 4);

  // ---- one-time setup -------------------------------------------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(wimg + dir * img_stride);
    float4* dst = reinterpret_cast<float4*>(s_wx);
    constexpr int n4 = (WX_BYTES + WH_BYTES) / 16;
    for (int i = tid; i < n4; i += NTHREADS) dst[i] = __ldg(src + i);
    const float* bsrc = wimg + dir * img_stride + (WX_BYTES + WH_BYTES) / 4;
    for (int i = tid; i < N4; i += NTHREADS) s_bias[i] = __ldg(bsrc + i);
  }
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(BAR_FULL(s), 1);
      mbar_init(BAR_EMPTY(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(BAR_ACC_FULL(b), 1);
      mbar_init(BAR_ACC_EMPTY(b), 8);
    }
    mbar_init(BAR_H, 8);
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
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      const uint32_t wx = smem_u32(s_wx), wh = smem_u32(s_wh), hs = smem_u32(s_h);
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
          const uint32_t xa = smem_u32(s_x + stage * SB_BYTES);
          for (int j = 0; j < chunks / 2; ++j) {
            const uint64_t ad = make_desc(xa + j * 2 * CHUNK_BYTES_A, CHUNK_BYTES_A, 128);
            const uint64_t bd = make_desc(wx + (sb * SB_CHUNKS + 2 * j) * CHUNK_BYTES_W, CHUNK_BYTES_W, 128);
            mma_tf32_ss(tacc, ad, bd, kIdesc, (sb | j) ? 1u : 0u);
          }
          mma_commit(BAR_EMPTY(stage));  // frees the ring stage once these MMAs have read it
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        if (step > 0) {
          if (!mbar_wait(BAR_H, (uint32_t)((step - 1) & 1), s_abort, gerr, 4)) break;
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < KH_CHUNKS / 2; ++j) {
            const uint64_t ad = make_desc(hs + j * 2 * CHUNK_BYTES_A, CHUNK_BYTES_A, 128);
            const uint64_t bd = make_desc(wh + j * 2 * CHUNK_BYTES_W, CHUNK_BYTES_W, 128);
            mma_tf32_ss(tacc, ad, bd, kIdesc, 1u);
          }
        }
        mma_commit(BAR_ACC_FULL(b));
      }
    }
  } else {
    // ===== epilogue: 8 warps; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    float c[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) c[i] = 0.f;
    float4* hs4 = reinterpret_cast<float4*>(s_h);
    for (int step = 0; step < T; ++step) {
      const int t = dir ? T - 1 - step : step;
      const int b = step & 1;
      if (!mbar_wait(BAR_ACC_FULL(b), (uint32_t)((step >> 1) & 1), s_abort, gerr, 5)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * ACC_COLS + half * 96);
      float4* hg4 = reinterpret_cast<float4*>(hout) +
                    (((int64_t)t * ntiles + tile) * (2 * KH_CHUNKS) + dir * KH_CHUNKS) * (TM) + row;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float v[32];
        tmem_ld32(taddr + ch * 32, v);
        if (ch == 2) {  // accumulator fully read by this warp: hand the buffer back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR_ACC_EMPTY(b));
        }
        float hv[8], hraw[8];
        float4* gc4 = nullptr;
        if (STASH)
          gc4 = reinterpret_cast<float4*>(gc) + ((((int64_t)dir * T + t) * ntiles + tile) * GC_CHUNKS) * TM + row;
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          const int ul = ch * 8 + uu;                 // unit index inside this thread's 24
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + (half * 24 + ul) * 4);
          const float ig = sigmoid_fast(v[4 * uu + 0] + bb.x);
          const float fg = sigmoid_fast(v[4 * uu + 1] + bb.y);
          const float gg = tanh_fast(v[4 * uu + 2] + bb.z);
          const float og = sigmoid_fast(v[4 * uu + 3] + bb.w);
          c[ul] = fg * c[ul] + ig * gg;
          hraw[uu] = og * tanh_fast(c[ul]);
          hv[uu] = rna_tf32(hraw[uu]);
          if (STASH) gc4[(int64_t)(half * 24 + ul) * TM] = make_float4(ig, fg, gg, og);
        }
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
          const int chunk = half * 6 + ch * 2 + k2;   // K chunk (4 hidden units) of this direction's h
          const float4 q = make_float4(hv[4 * k2], hv[4 * k2 + 1], hv[4 * k2 + 2], hv[4 * k2 + 3]);
          hs4[chunk * TM + row] = q;                  // next step's A operand
          hg4[(int64_t)chunk * TM] = q;               // next layer's input (coalesced: lane = row)
          if (STASH) {
            const int ul = ch * 8 + 4 * k2;
            gc4[(int64_t)(HID + chunk) * TM] = make_float4(c[ul], c[ul + 1], c[ul + 2], c[ul + 3]);
            const int64_t bidx = (int64_t)tile * TM + row;
            if (bidx < B)
              *reinterpret_cast<float4*>(h_rm + ((int64_t)t * B + bidx) * (2 * HID) + dir * HID + chunk * 4) =
                  make_float4(hraw[4 * k2], hraw[4 * k2 + 1], hraw[4 * k2 + 2], hraw[4 * k2 + 3]);
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_H);
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
//   gc     : [2][T][ntiles][60][128][4]   (forward stash)         dh_out : [T][B][2H] row-major
//   da_rm  : [2][T][B][4H] row-major, PyTorch gate order (i,f,g,o blocks of H)
// ---------------------------------------------------------------------------------------------
constexpr int BWD_THREADS = 288;  // warp 0: MMA issuer; warps 1..8: epilogue (TMEM quarter = warp % 4)
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
                                                                     float* __restrict__ da_rm, int T, int ntiles,
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

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TM, HID);
      const uint32_t da = smem_u32(s_da), wb = smem_u32(s_w);
      int n = 0;
      for (int step = T - 1; step >= 1; --step, ++n) {
        if (!mbar_wait(BAR_DA, (uint32_t)(n & 1), s_abort, gerr, 31)) break;
        tc_fence_after();
#pragma unroll 4
        for (int j = 0; j < N4 / 8; ++j) {
          const uint64_t ad = make_desc(da + j * 2 * CHUNK_BYTES_A, CHUNK_BYTES_A, 128);
          const uint64_t bd = make_desc(wb + j * 2 * WT_CHUNK_BYTES, WT_CHUNK_BYTES, 128);
          mma_tf32_ss(tmem_base, ad, bd, idesc, j ? 1u : 0u);
        }
        mma_commit(BAR_ACC);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 1) >> 2;
    const int row = quarter * 32 + lane;
    const int64_t bidx = (int64_t)tile * TM + row;
    const bool valid = bidx < B;
    float dc[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) dc[i] = 0.f;
    float4* da4 = reinterpret_cast<float4*>(s_da);
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 24);
    int n = 0;
    for (int step = T - 1; step >= 0; --step, ++n) {
      const int t = dir ? T - 1 - step : step;
      const int tp = dir ? t + 1 : t - 1;
      const float4* g4 = reinterpret_cast<const float4*>(gc) + ((((int64_t)dir * T + t) * ntiles + tile) * GC_CHUNKS) * TM + row;
      const float4* gp4 = reinterpret_cast<const float4*>(gc) + ((((int64_t)dir * T + tp) * ntiles + tile) * GC_CHUNKS) * TM + row;
      const float* dhp = dh_out + ((int64_t)t * B + bidx) * (2 * HID) + dir * HID + half * 24;
      float* dap = da_rm + (((int64_t)dir * T + t) * B + bidx) * N4 + half * 24;
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
        float dho[8], cc[8], cp[8];
        {
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 a0 = valid ? __ldg(reinterpret_cast<const float4*>(dhp + ch * 8)) : z4;
          const float4 a1 = valid ? __ldg(reinterpret_cast<const float4*>(dhp + ch * 8 + 4)) : z4;
          dho[0] = a0.x; dho[1] = a0.y; dho[2] = a0.z; dho[3] = a0.w; dho[4] = a1.x; dho[5] = a1.y; dho[6] = a1.z; dho[7] = a1.w;
          const int cchunk = HID + half * 6 + ch * 2;
          const float4 c0 = __ldg(g4 + (int64_t)cchunk * TM), c1 = __ldg(g4 + (int64_t)(cchunk + 1) * TM);
          cc[0] = c0.x; cc[1] = c0.y; cc[2] = c0.z; cc[3] = c0.w; cc[4] = c1.x; cc[5] = c1.y; cc[6] = c1.z; cc[7] = c1.w;
          if (step > 0) {
            const float4 p0 = __ldg(gp4 + (int64_t)cchunk * TM), p1 = __ldg(gp4 + (int64_t)(cchunk + 1) * TM);
            cp[0] = p0.x; cp[1] = p0.y; cp[2] = p0.z; cp[3] = p0.w; cp[4] = p1.x; cp[5] = p1.y; cp[6] = p1.z; cp[7] = p1.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) cp[i] = 0.f;
          }
        }
        float dai[8], daf[8], dag[8], dao[8];
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          const int ul = ch * 8 + uu;
          const float4 gt = __ldg(g4 + (int64_t)(half * 24 + ul) * TM);  // (i, f, g, o)
          const float tch = tanh_fast(cc[uu]);
          const float dh = dho[uu] + rec[uu];
          const float d_o = dh * tch;
          const float dct = dc[ul] + dh * gt.w * (1.f - tch * tch);
          dai[uu] = dct * gt.z * gt.x * (1.f - gt.x);
          daf[uu] = dct * cp[uu] * gt.y * (1.f - gt.y);
          dag[uu] = dct * gt.x * (1.f - gt.z * gt.z);
          dao[uu] = d_o * gt.w * (1.f - gt.w);
          dc[ul] = dct * gt.y;
          da4[(half * 24 + ul) * TM + row] = make_float4(rna_tf32(dai[uu]), rna_tf32(daf[uu]), rna_tf32(dag[uu]), rna_tf32(dao[uu]));
        }
        if (valid) {
          float4* o;
          o = reinterpret_cast<float4*>(dap + 0 * HID + ch * 8);
          o[0] = make_float4(dai[0], dai[1], dai[2], dai[3]); o[1] = make_float4(dai[4], dai[5], dai[6], dai[7]);
          o = reinterpret_cast<float4*>(dap + 1 * HID + ch * 8);
          o[0] = make_float4(daf[0], daf[1], daf[2], daf[3]); o[1] = make_float4(daf[4], daf[5], daf[6], daf[7]);
          o = reinterpret_cast<float4*>(dap + 2 * HID + ch * 8);
          o[0] = make_float4(dag[0], dag[1], dag[2], dag[3]); o[1] = make_float4(dag[4], dag[5], dag[6], dag[7]);
          o = reinterpret_cast<float4*>(dap + 3 * HID + ch * 8);
          o[0] = make_float4(dao[0], dao[1], dao[2], dao[3]); o[1] = make_float4(dao[4], dao[5], dao[6], dao[7]);
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR_DA);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
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

template <int KXC, int STASH>
int launch_layer(wgg_ctx* ctx, const float* xin, const float* img, int64_t img_stride, float* hout, int T, int ntiles,
                 int64_t B, float* gc, float* h_rm, cudaStream_t st) {
  constexpr size_t smem = (size_t)KXC * tc::CHUNK_BYTES_W + tc::KH_CHUNKS * tc::CHUNK_BYTES_W + tc::NSTAGE * tc::SB_BYTES +
                          tc::KH_CHUNKS * tc::CHUNK_BYTES_A + tc::N4 * 4 + 16 * 8 + 16;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tc::lstm_tc_fwd_kernel<KXC, STASH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return wgg_fail(ctx, WGG_ECUDA, "lstm_tc_fwd_kernel: cannot reserve shared memory%s");
    configured = true;
  }
  dim3 grid((unsigned)ntiles, 2);
  // algorithmic FLOPs: 2 dirs x T x B x 2 x 192 x (K_x + 48); bytes: x in (both dirs read it) + h out
  ProfScope prof(ctx, "lstm_tc_fwd_kernel", st, 2.0 * T * (double)B * 2.0 * tc::N4 * (KXC * 4 + tc::HID),
                 (double)T * B * 4.0 * (2.0 * KXC * 4 + 96));
  tc::lstm_tc_fwd_kernel<KXC, STASH><<<grid, tc::NTHREADS, smem, st>>>(xin, img, img_stride, hout, T, ntiles, gc, h_rm, B,
                                                                         ctx->async_err);
  WGG_CHECK_LAUNCH(ctx, "lstm_tc_fwd_kernel");
  return WGG_OK;
}
}  // namespace

int64_t generator_tc_gc_layer_floats(const wgg_model_cfg* cfg, int64_t B) {
  TcPlan p;
  if (!tc_plan(cfg, B, &p)) return 0;
  return (int64_t)2 * p.T * p.rows * (tc::GC_CHUNKS * 4);
}

// BPTT of one layer (both directions) on the tcgen05 path: gc (forward stash) + dh_out -> da_rm
int lstm_tc_bwd_layer(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* lp, int64_t dir_stride, int64_t off_whh,
                      const float* gc, const float* dh_out, float* da_rm, float* wimg_ws, int64_t B, cudaStream_t st) {
  TcPlan p;
  if (!tc_plan(cfg, B, &p)) return wgg_fail(ctx, WGG_EUNSUPPORTED, "lstm_tc_bwd_layer: unsupported configuration%s");
  tc::prep_whhT_kernel<<<dim3(8, 2), 256, 0, st>>>(lp, dir_stride, off_whh, wimg_ws);
  WGG_CHECK_LAUNCH(ctx, "prep_whhT_kernel");
  constexpr size_t smem = (size_t)tc::HID * tc::CHUNK_BYTES_A + tc::HID * tc::WT_CHUNK_BYTES + 4 * 8 + 16;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tc::lstm_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return wgg_fail(ctx, WGG_ECUDA, "lstm_tc_bwd_kernel: cannot reserve shared memory%s");
    configured = true;
  }
  dim3 grid((unsigned)p.ntiles, 2);
  ProfScope prof(ctx, "lstm_tc_bwd_kernel", st, 2.0 * p.T * (double)B * 2.0 * tc::N4 * tc::HID,
                 2.0 * p.T * (double)B * 4.0 * (tc::GC_CHUNKS * 4 + tc::HID + tc::HID + tc::N4));
  tc::lstm_tc_bwd_kernel<<<grid, tc::BWD_THREADS, smem, st>>>(gc, wimg_ws, dh_out, da_rm, p.T, p.ntiles, B, ctx->async_err);
  WGG_CHECK_LAUNCH(ctx, "lstm_tc_bwd_kernel");
  return WGG_OK;
}

int64_t generator_tc_workspace_floats(const wgg_model_cfg* cfg, int64_t B) {
  TcPlan p;
  return tc_plan(cfg, B, &p) ? p.total : 0;
}

bool generator_tc_supported(const wgg_model_cfg* cfg) {
  TcPlan p;
  return tc_plan(cfg, 1, &p);
}

// no-grad generator forward on the tcgen05 path.  `layer_off`, `dir_stride`, ... describe the flat parameter layout.
int generator_forward_tc(wgg_ctx* ctx, const wgg_model_cfg* cfg, const float* params, const int64_t* layer_off,
                         const int64_t* dir_stride, const int64_t* off_whh, const int64_t* off_bih,
                         const int64_t* off_bhh, int64_t off_wo, int64_t off_bo, const float* proto, const float* z,
                         int64_t B, float* out, float* ws, int64_t ws_floats, float* gc_stash, float* const* hseq_rm,
                         cudaStream_t st) {
  TcPlan p;
  const int64_t gc_layer_floats = generator_tc_gc_layer_floats(cfg, B);
  if (!tc_plan(cfg, B, &p)) return wgg_fail(ctx, WGG_EUNSUPPORTED, "generator_forward_tc: unsupported configuration%s");
  if (!ws || ws_floats < p.total) return wgg_fail(ctx, WGG_EWORKSPACE, "generator_forward_tc: workspace too small%s");
  float* x0 = ws;
  float* hbuf[2] = {ws + p.x0_floats, ws + p.x0_floats + p.h_floats};
  for (int l = 0; l < p.L; ++l) {
    const int I = l == 0 ? p.pd + p.Z : 96;
    const int KX = l == 0 ? kKX0 : 96;
    tc::prep_weights_kernel<<<dim3(32, 2), 256, 0, st>>>(params + layer_off[l], dir_stride[l], off_whh[l], off_bih[l],
                                                         off_bhh[l], I, KX, ws + p.img_off[l], p.img_floats[l]);
    WGG_CHECK_LAUNCH(ctx, "prep_weights_kernel");
  }
  tc::build_x0_tc_kernel<<<ew_blocks(p.x0_floats), 256, 0, st>>>(proto, z, x0, p.T, B, p.ntiles, p.C, p.pd, p.Z, kKX0);
  WGG_CHECK_LAUNCH(ctx, "build_x0_tc_kernel");
  const float* in = x0;
  for (int l = 0; l < p.L; ++l) {
    float* hout = hbuf[l & 1];
    float* gcl = gc_stash ? gc_stash + (int64_t)l * gc_layer_floats : nullptr;
    float* hrm = gc_stash ? hseq_rm[l] : nullptr;
    if (l == 0) {
      if (gc_stash) WGG_TRY((launch_layer<kKX0 / 4, 1>(ctx, in, ws + p.img_off[l], p.img_floats[l], hout, p.T, p.ntiles, B, gcl, hrm, st)));
      else WGG_TRY((launch_layer<kKX0 / 4, 0>(ctx, in, ws + p.img_off[l], p.img_floats[l], hout, p.T, p.ntiles, B, nullptr, nullptr, st)));
    } else {
      if (gc_stash) WGG_TRY((launch_layer<24, 1>(ctx, in, ws + p.img_off[l], p.img_floats[l], hout, p.T, p.ntiles, B, gcl, hrm, st)));
      else WGG_TRY((launch_layer<24, 0>(ctx, in, ws + p.img_off[l], p.img_floats[l], hout, p.T, p.ntiles, B, nullptr, nullptr, st)));
    }
    in = hout;
  }
  dim3 grid((unsigned)cdiv64(B, 32), (unsigned)((p.T + 7) / 8));
  tc::head_tc_kernel<<<grid, 256, 0, st>>>(in, params + off_wo, params + off_bo, out, p.T, B, p.ntiles, p.C);
  WGG_CHECK_LAUNCH(ctx, "head_tc_kernel");
  return WGG_OK;
}

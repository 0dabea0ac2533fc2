// Micro-probe of tcgen05.mma operand / accumulator conventions (M=64/128, K-major / MN-major, no swizzle).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../wordgesture-gan_b200/csrc/tc_common.cuh"
using namespace tcu;

struct Cfg { int M, N, a_mn, b_mn, swap_a, swap_b; };

// smem layouts (floats): K-major operand rows x 8: element (r,k): (k/4)*CSK + r*4 + k%4  with CSK = rows*4 floats
//                        MN-major operand rows x 8: element (r,k): (r/4)*CSM + k*4 + r%4  with CSM = 8*4*... we use CSM = 64 floats (=256B) spacing
__global__ void probe(Cfg c, float* out /*128 x 256*/, int* gerr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sa = reinterpret_cast<float*>(smem);             // 16 KB
  float* sb = reinterpret_cast<float*>(smem + 16384);     // 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 8192; i += blockDim.x) { sa[i] = 0.f; }
  for (int i = tid; i < 4096; i += blockDim.x) { sb[i] = 0.f; }
  __syncthreads();
  const int CSM = 64;  // floats between MN chunks in MN-major tiles
  for (int i = tid; i < c.M * 8; i += blockDim.x) {
    const int m = i / 8, k = i % 8;
    const float v = (float)(m * 8 + k + 1);
    if (!c.a_mn) sa[(k / 4) * (c.M * 4) + m * 4 + (k % 4)] = v;
    else sa[(m / 4) * CSM + k * 4 + (m % 4)] = v;
  }
  for (int i = tid; i < c.N * 8; i += blockDim.x) {
    const int n = i / 8, k = i % 8;
    const float v = (k == (n % 8)) ? 1.f : 0.f;
    if (!c.b_mn) sb[(k / 4) * (c.N * 4) + n * 4 + (k % 4)] = v;
    else sb[(n / 4) * CSM + k * 4 + (n % 4)] = v;
  }
  if (tid == 0) { mbar_init(smem_u32(bar), 1); *s_abort = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *s_tmem;
  // clear TMEM via a zero MMA? simpler: read-before to know the background; we fill with an initial MMA using zero B
  if (tid == 0) {
    uint64_t ad, bd;
    if (!c.a_mn) ad = make_desc(smem_u32(sa), c.M * 16, 128);
    else ad = c.swap_a ? make_desc(smem_u32(sa), CSM * 4, 128) : make_desc(smem_u32(sa), 128, CSM * 4);
    if (!c.b_mn) bd = make_desc(smem_u32(sb), c.N * 16, 128);
    else bd = c.swap_b ? make_desc(smem_u32(sb), CSM * 4, 128) : make_desc(smem_u32(sb), 128, CSM * 4);
    const uint32_t idesc = make_idesc(c.M, c.N, c.a_mn, c.b_mn);
    mma_tf32_ss(tb, ad, bd, idesc, 0u);
    mma_commit(smem_u32(bar));
  }
  if (warp < 4) {
    mbar_wait(smem_u32(bar), 0, s_abort, gerr, 99);
    tc_fence_after();
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 256; c0 += 16) {
      float r[16];
      tmem_ld16(taddr + c0, r);
      for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 256 + c0 + i] = r[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

int main() {
  float* d_out; int* d_err;
  cudaMalloc(&d_out, 128 * 256 * 4); cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  Cfg cfgs[] = {{128, 16, 0, 0, 0, 0}, {64, 16, 0, 0, 0, 0}, {128, 16, 1, 0, 0, 0}, {128, 16, 1, 0, 1, 0}, {128, 16, 0, 1, 0, 0},
                {128, 16, 0, 1, 0, 1}, {64, 16, 1, 1, 0, 0}, {64, 16, 1, 1, 1, 1}, {64, 64, 1, 1, 0, 0}, {64, 8, 1, 1, 0, 0}};
  std::vector<float> h(128 * 256);
  for (auto& c : cfgs) {
    cudaMemset(d_out, 0, 128 * 256 * 4);
    probe<<<1, 128, 40960>>>(c, d_out, d_err);
    cudaError_t e = cudaDeviceSynchronize();
    int err = 0; cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h.data(), d_out, 128 * 256 * 4, cudaMemcpyDeviceToHost);
    printf("cfg M=%d N=%d a_mn=%d b_mn=%d swap_a=%d swap_b=%d : cuda=%s err=%d\n", c.M, c.N, c.a_mn, c.b_mn, c.swap_a, c.swap_b, cudaGetErrorString(e), err);
    // expected D[m][n] = A[m][n%8] = m*8 + n%8 + 1.  Report, for each lane, what "m" the values decode to.
    int good = 0, nonzero = 0;
    for (int lane = 0; lane < 128; ++lane) for (int n = 0; n < c.N; ++n) { float v = h[lane * 256 + n]; if (v != 0.f) ++nonzero; }
    printf("   nonzero entries in first N cols: %d\n", nonzero);
    for (int lane : {0, 1, 5, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 96, 127}) {
      printf("   lane %3d:", lane);
      for (int n = 0; n < (c.N < 12 ? c.N : 12); ++n) printf(" %6.0f", h[lane * 256 + n]);
      printf("\n");
    }
    (void)good;
  }
  return 0;
}

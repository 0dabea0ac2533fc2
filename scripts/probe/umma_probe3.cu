// Probe MN-major TF32 operands in the SWIZZLE_128B_BASE32B layout (the only MN-major layout tf32 supports),
// including start addresses shifted by whole K rows (sliding window) and M=64.
#include <cstdio>
#include <vector>
#include "../../wordgesture-gan_b200/csrc/tc_common.cuh"
using namespace tcu;

struct Cfg { int M, N, a_mn, b_mn, shift, swap, col, blk_rows, row0; };

__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// MN-major SW128_32B tile: element (mn, k) at byte  (mn/32)*BLK + k*128 + (((mn%32)/8) ^ (k&3))*32 + (mn%8)*4
// (row = k, 128 B per row = 32 MN elements, 32-byte pieces XOR-swizzled by the row index; tile base 1024-aligned)
__device__ __forceinline__ int mn_sw_off(int mn, int k, int BLK) {
  return (mn / 32) * BLK + k * 128 + ((((mn % 32) / 8) ^ (k & 3)) * 32) + (mn % 8) * 4;
}

__global__ void probe(Cfg c, float* out, int* gerr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;             // 32 KB
  uint8_t* sb = smem + 40960;     // 40 KB each
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 81920);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* s_abort = reinterpret_cast<volatile int*>(s_tmem + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 20480; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  const int KROWS = 32, BLK = c.blk_rows * 128;  // each 32-wide MN block holds blk_rows K rows (32 used)
  for (int i = tid; i < c.M * KROWS; i += blockDim.x) {
    const int m = i / KROWS, k = i % KROWS;
    const float v = (float)(m * 16 + (k % 16) + 1);
    if (!c.a_mn) { if (k < 8) reinterpret_cast<float*>(sa)[(k / 4) * (c.M * 4) + m * 4 + (k % 4)] = v; }
    else *reinterpret_cast<float*>(sa + mn_sw_off(m, k + c.row0, BLK)) = v;
  }
  for (int i = tid; i < c.N * KROWS; i += blockDim.x) {
    const int n = i / KROWS, k = i % KROWS;
    const float v = (k == (n % 8) + c.shift * (c.b_mn ? 1 : 0)) ? 1.f : 0.f;   // identity on the (shifted) window
    if (!c.b_mn) { if (k < 8) reinterpret_cast<float*>(sb)[(k / 4) * (c.N * 4) + n * 4 + (k % 4)] = (k == n % 8) ? 1.f : 0.f; }
    else *reinterpret_cast<float*>(sb + mn_sw_off(n, k + c.row0, BLK)) = v;
  }
  if (tid == 0) { mbar_init(smem_u32(bar), 1); *s_abort = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *s_tmem;
  if (tid == 0) {
    uint64_t ad, bd;
    const uint32_t sh = (c.shift + c.row0) * 128;
    if (!c.a_mn) ad = make_desc(smem_u32(sa), c.M * 16, 128);
    else ad = c.swap ? make_desc_sw(smem_u32(sa) + sh, 512, BLK, 1) : make_desc_sw(smem_u32(sa) + sh, BLK, 512, 1);
    if (!c.b_mn) bd = make_desc(smem_u32(sb), c.N * 16, 128);
    else bd = c.swap ? make_desc_sw(smem_u32(sb) + sh, 512, BLK, 1) : make_desc_sw(smem_u32(sb) + sh, BLK, 512, 1);
    mma_tf32_ss(tb + c.col, ad, bd, make_idesc(c.M, c.N, c.a_mn, c.b_mn), 0u);
    mma_commit(smem_u32(bar));
  }
  if (warp < 4) {
    mbar_wait(smem_u32(bar), 0, s_abort, gerr, 99);
    tc_fence_after();
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 16) {
      float r[16];
      tmem_ld16(taddr + c.col + c0, r);
      for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c0 + i] = r[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  float* d_out; int* d_err;
  cudaMalloc(&d_out, 128 * 64 * 4); cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 82432);
  Cfg cfgs[] = {{64, 64, 1, 1, 0, 0, 0, 32, 0}, {64, 64, 1, 1, 0, 0, 320, 32, 0}, {64, 64, 1, 1, 0, 0, 0, 136, 0}, {64, 64, 1, 1, 0, 0, 0, 32, 2}, {64, 64, 1, 1, 1, 0, 128, 136, 42}, {64, 8, 1, 1, 0, 0, 320, 136, 42}, {128, 64, 1, 1, 0, 0, 64, 136, 10}};
  std::vector<float> h(128 * 64);
  for (auto& c : cfgs) {
    cudaMemset(d_out, 0, 128 * 64 * 4);
    probe<<<1, 128, 82432>>>(c, d_out, d_err);
    cudaError_t e = cudaDeviceSynchronize();
    int err = 0; cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h.data(), d_out, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    // expectation: D[m][n] = A[m][(n%8) + shift_a] = m*16 + (n%8) + shift_a + 1  (shift_a = shift if A is MN-major;
    // if only B is MN-major the B identity is built on the shifted window so D[m][n] = A[m][n%8])
    const int sa = c.a_mn ? c.shift : 0;
    int ok = 0, tot = 0, nz = 0;
    for (int m = 0; m < c.M; ++m) {
      const int lane = c.M == 128 ? m : (m % 16) + 32 * (m / 16);
      for (int n = 0; n < c.N; ++n) {
        const float v = h[lane * 64 + n];
        // when both are shifted windows (a_mn && b_mn): A row k+shift pairs with B row k+shift -> D = A[m][(n%8)+shift]
        const float exp = (float)(m * 16 + ((n % 8) + sa) % 16 + 1);
        ++tot; if (v == exp) ++ok; if (v != 0.f) ++nz;
      }
    }
    printf("cfg M=%d N=%d shift=%d col=%d blk_rows=%d row0=%d : cuda=%s err=%d  match %d/%d nonzero %d\n", c.M, c.N, c.shift, c.col, c.blk_rows, c.row0, cudaGetErrorString(e), err, ok, tot, nz);
    for (int lane : {0, 1, 33, 127}) {
      printf("   lane %3d:", lane);
      for (int n = 0; n < 10; ++n) printf(" %6.0f", h[lane * 64 + n]);
      printf("\n");
    }
  }
  return 0;
}

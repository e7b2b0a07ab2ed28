// Probe for MN-major operands: D[128 n x 128 k] = sum_r G[r][n] * X[r][k], r < R (multiple of 32), both operands stored
// as natural row-major [r][feature] chunks (4 MN-blocks of [32 r][32 f], 128-byte rows, SWIZZLE_128B).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#define CHUNK_BYTES (32 * 128 * 4)   // [32 r][128 f] fp32 = 16 KB = 4 blocks of 4 KB
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((4096 >> 4) & 0x3FFF) << 16;   // LBO: next 32-wide MN block
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;    // SBO: next atom of 4 K-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                        // SWIZZLE_128B_BASE32B (the only MN-major layout for tf32)
  return d;
}
// element (row r in chunk, feature f): block f/32, row r, 16B chunk (f%32)/4 swizzled with r%8
__device__ __forceinline__ uint32_t mn_off(int r, int f) {
  const int c = (f & 31) >> 3;   // 32-byte chunk inside the 128-byte row
  return (uint32_t)((f >> 5) * 4096 + r * 128 + ((c ^ (r & 3)) << 5) + (f & 7) * 4);
}
__global__ void __launch_bounds__(256, 1) probe(const float* G, const float* X, float* D, int R, int split) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nch = R / 32;
  uint8_t* g_hi = smem; uint8_t* g_lo = g_hi + nch * CHUNK_BYTES;
  uint8_t* x_hi = g_lo + nch * CHUNK_BYTES; uint8_t* x_lo = x_hi + nch * CHUNK_BYTES;
  __shared__ uint64_t bar; __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < R * 128; idx += blockDim.x) {
    const int r = idx / 128, f = idx % 128;
    const float gv = G[idx], xv = X[idx];
    uint32_t u; float gh, gl, xh, xl;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(gv)); gh = __uint_as_float(u); gl = gv - gh;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(xv)); xh = __uint_as_float(u); xl = xv - xh;
    const uint32_t off = (r / 32) * CHUNK_BYTES + mn_off(r % 32, f);
    *reinterpret_cast<float*>(g_hi + off) = split ? gh : gv; *reinterpret_cast<float*>(g_lo + off) = gl;
    *reinterpret_cast<float*>(x_hi + off) = split ? xh : xv; *reinterpret_cast<float*>(x_lo + off) = xl;
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int first = 1;
    for (int p = 0; p < (split ? 3 : 1); ++p) {
      const uint8_t* pa = (p == 2) ? g_lo : g_hi; const uint8_t* pb = (p == 1) ? x_lo : x_hi;
      for (int ch = 0; ch < nch; ++ch)
        for (int j = 0; j < 4; ++j) {   // 4 groups of 8 rows
          const uint64_t da = make_desc_mn(smem_u32(pa + ch * CHUNK_BYTES) + j * 1024);
          const uint64_t db = make_desc_mn(smem_u32(pb + ch * CHUNK_BYTES) + j * 1024);
          const uint32_t acc = first ? 0u : 1u; first = 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  { uint32_t ok = 0; while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                     "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                     "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(size_t)row * 128 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}
int main() {
  for (int R : {32, 64}) for (int split = 0; split < 2; ++split) {
    std::vector<float> G(R * 128), X(R * 128), D(128 * 128, -1.f);
    srand(2);
    for (auto& x : G) x = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& x : X) x = (float)rand() / RAND_MAX * 2 - 1;
    float *dG, *dX, *dD;
    cudaMalloc(&dG, G.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    size_t smem = (size_t)4 * (R / 32) * CHUNK_BYTES + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 256, smem>>>(dG, dX, dD, R, split);
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
    printf("R=%d split=%d: %s\n", R, split, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int n = 0; n < 128; ++n) for (int k = 0; k < 128; ++k) {
      double r = 0; for (int i = 0; i < R; ++i) r += (double)G[i * 128 + n] * (double)X[i * 128 + k];
      maxerr = fmax(maxerr, fabs(r - D[n * 128 + k])); maxref = fmax(maxref, fabs(r));
    }
    printf("   max|ref| %.4f  max err %.3e  rel %.3e\n", maxref, maxerr, maxerr / maxref);
  }
  return 0;
}

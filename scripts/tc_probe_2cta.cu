// Standalone probe for round 2's CTA-pair tensor-core kernels: a 2-CTA cluster computes
//     D[256 x 128] = A[256 x K] * B[128 x K]^T          (tf32, 1x and 3x split)
// with ONE tcgen05.mma.cta_group::2 stream issued by the leader CTA: UMMA M = 256 (each CTA supplies its own 128 rows
// of A), N = 128 split across the pair (each CTA holds 64 of the 128 rows of B, i.e. half of the weight).  Checks the
// operand placement, the instruction descriptor, the multicast commit and the per-CTA accumulator layout against an
// fp64 host product.  Every wait is bounded (trap instead of hanging the GPU).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define MT 256          // rows of D for the pair
#define N 128
#define KB 32           // floats per K-block (128 B rows)
#define A_BLK (128 * 128)   // [128 rows x 32 floats]
#define B_BLK (64 * 128)    // [ 64 rows x 32 floats]

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;   // SBO = 1024 B between 8-row groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bounded_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 2000000000ll) { printf("probe: mbarrier wait timed out (cta %d)\n", (int)blockIdx.x); __trap(); }
  }
}
#define SPLIT(x, h, l) { h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); l = x - h; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) probe2(const float* A, const float* B, float* D, int K, int split) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int nkb = K / KB;
  uint8_t* a_hi = smem;                        // [nkb][16 KB]  this CTA's 128 rows of A
  uint8_t* a_lo = a_hi + nkb * A_BLK;
  uint8_t* b_hi = a_lo + nkb * A_BLK;          // [nkb][ 8 KB]  this CTA's 64 rows of B
  uint8_t* b_lo = b_hi + nkb * B_BLK;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int idx = tid; idx < 128 * (K / 4); idx += blockDim.x) {
    const int r = idx / (K / 4), c4 = idx % (K / 4), kb = c4 / 8, c = c4 % 8;
    const float4 va = *reinterpret_cast<const float4*>(A + (size_t)(rank * 128 + r) * K + c4 * 4);
    float4 h, l;
    SPLIT(va.x, h.x, l.x) SPLIT(va.y, h.y, l.y) SPLIT(va.z, h.z, l.z) SPLIT(va.w, h.w, l.w)
    const uint32_t off = kb * A_BLK + sw128_off(r, c);
    *reinterpret_cast<float4*>(a_hi + off) = split ? h : va;
    *reinterpret_cast<float4*>(a_lo + off) = l;
  }
  for (int idx = tid; idx < 64 * (K / 4); idx += blockDim.x) {
    const int r = idx / (K / 4), c4 = idx % (K / 4), kb = c4 / 8, c = c4 % 8;
    const float4 vb = *reinterpret_cast<const float4*>(B + (size_t)(rank * 64 + r) * K + c4 * 4);
    float4 h, l;
    SPLIT(vb.x, h.x, l.x) SPLIT(vb.y, h.y, l.y) SPLIT(vb.z, h.z, l.z) SPLIT(vb.w, h.w, l.w)
    const uint32_t off = kb * B_BLK + sw128_off(r, c);
    *reinterpret_cast<float4*>(b_hi + off) = split ? h : vb;
    *reinterpret_cast<float4*>(b_lo + off) = l;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();   // both CTAs: operands written, barriers initialised, TMEM allocated
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (rank == 0 && tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
    int first = 1;
    const int npass = split ? 3 : 1;
    for (int p = 0; p < npass; ++p) {
      const uint8_t* pa = (p == 2) ? a_lo : a_hi;   // passes: hi*hi, hi*lo, lo*hi
      const uint8_t* pb = (p == 1) ? b_lo : b_hi;
      for (int kb = 0; kb < nkb; ++kb)
        for (int j = 0; j < 4; ++j) {
          const uint64_t da = make_desc(smem_u32(pa + kb * A_BLK) + j * 32);
          const uint64_t db = make_desc(smem_u32(pb + kb * B_BLK) + j * 32);
          const uint32_t acc = first ? 0u : 1u;
          first = 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  bounded_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    const int row = rank * 128 + warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                     "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                     "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();   // neither CTA may free tensor memory while the other still reads it
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
  for (int K : {32, 128}) {
    for (int split = 0; split < 2; ++split) {
      std::vector<float> A(MT * K), B(N * K), D(MT * N, -1.f);
      srand(1);
      for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
      for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
      float *dA, *dB, *dD;
      cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
      cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
      cudaMemset(dD, 0xff, D.size() * 4);
      size_t smem = (size_t)2 * (K / KB) * (A_BLK + B_BLK) + 1024;
      cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      probe2<<<2, 256, smem>>>(dA, dB, dD, K, split);
      cudaError_t e = cudaDeviceSynchronize();
      printf("2cta K=%d split=%d: %s\n", K, split, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int i = 0; i < MT; ++i)
        for (int j = 0; j < N; ++j) {
          double r = 0;
          for (int k = 0; k < K; ++k) r += (double)A[i * K + k] * (double)B[j * K + k];
          maxerr = fmax(maxerr, fabs(r - D[i * N + j]));
          maxref = fmax(maxref, fabs(r));
        }
      printf("   max|ref| %.4f  max err %.3e  rel %.3e\n", maxref, maxerr, maxerr / maxref);
      cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
  }
  return 0;
}

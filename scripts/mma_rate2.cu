// (two-issuer variant of mma_rate.cu: do two warps issuing to different accumulators share the >= 105-cycle floor?)
// How long does ONE tcgen05.mma kind::tf32 (M = 128, K = 8) take on a B200 SM, as a function of N, of where the A operand
// lives (shared memory vs tensor memory) and of whether consecutive instructions share an accumulator?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/mma_rate2 scripts/mma_rate2.cu && scripts/build/mma_rate2
// One thread per CTA issues L instructions back to back, commits, waits; cycles = clock64 difference / L.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// whole warp converged: the election happens inside the asm block, the MMA is predicated on it
__device__ __forceinline__ void mma_ss_warp(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c));
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// mode 0: SS same accumulator; 1: TS same accumulator; 2: SS two accumulators alternating; 3: TS two accumulators
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int L, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) f[i] = 1.0f + (i % 7) * 0.125f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  __shared__ uint64_t bar2[4];
  __shared__ long long tt[4];
  if (threadIdx.x == 0) { for (int w = 0; w < 4; ++w) mbar_init(&bar2[w], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const bool whole_warp = mode >= 8;
  const int issuers = mode & 7;   // 1, 2 or 4 warps issue L / issuers instructions each, each to its own accumulator
  const int w = threadIdx.x >> 5;
  if ((whole_warp || (threadIdx.x & 31) == 0) && w < issuers) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    uint32_t parity = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < L / issuers; ++i) {
        const uint32_t o = (i & 3) * 32;
        if (whole_warp) mma_ss_warp(tmem + (uint32_t)w * 128u, make_desc(sa + o), make_desc(sb + o), idesc, i >= 1 ? 1u : 0u);
        else mma_ss(tmem + (uint32_t)w * 128u, make_desc(sa + o), make_desc(sb + o), idesc, i >= 1 ? 1u : 0u);
      }
      if ((threadIdx.x & 31) == 0)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[w])) : "memory");
      while (!try_wait(&bar2[w], parity)) {}
      parity ^= 1u;
      const long long t1 = clock64();
      if (rep == 2 && (threadIdx.x & 31) == 0) tt[w] = t1 - t0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long m = 0;
    for (int q = 0; q < issuers; ++q) m = tt[q] > m ? tt[q] : m;
    out[0] = m;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 16384 + 32768;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int L = 960;
  for (int grid : {1, 148}) {
    printf("grid %d CTAs, %d instructions in total (M = 128, K = 8, kind::tf32, operands in shared memory)\n", grid, L);
    for (int issuers : {1, 2, 4, 9, 10}) {
      for (int N : {64, 128}) {
        rate_kernel<<<grid, 128, smem>>>(N, L, issuers, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("issuers %d N=%d: %s\n", issuers, N, cudaGetErrorString(e)); return 1; }
        long long c;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        const double per = (double)c / L;
        printf("  %d issuing warp(s)%s  N=%3d : %7.1f cycles per MMA  (%.0f MAC/clk; dense tf32 peak ~1960)\n", issuers & 7,
               issuers >= 8 ? ", converged warp + elect.sync" : "", N, per,
               128.0 * N * 8 / per);
      }
    }
  }
  return 0;
}

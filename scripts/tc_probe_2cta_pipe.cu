// Pipeline skeleton of round 2's CTA-pair Linear kernel, as a standalone probe:
//     Y[R x 128] = X[R x 128] * W[128 x 128]^T   (3xTF32), R = T * 256 rows, several 2-CTA clusters, persistent over tiles.
// Per CTA: 8 producer warps split their 128 rows of X into a 2-stage operand ring, the leader CTA's MMA thread issues
// tcgen05.mma.cta_group::2 (UMMA M = 256; each CTA keeps only its 64 weight rows resident), 4 epilogue warps drain the
// CTA's own accumulator (2 TMEM buffers).  What this validates beyond tc_probe_2cta.cu:
//   * full[stage] / acc_free[buf] live in the LEADER and collect arrivals from both CTAs (mapa + cluster-scope arrive),
//   * mma_done[stage] / acc_done[buf] are signalled in BOTH CTAs by one multicast tcgen05.commit,
//   * ring + accumulator recycling across tiles with those cross-CTA barriers.
// Every wait is bounded (trap instead of hanging the GPU).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define KD 128
#define N 128
#define KB 32
#define NKB (KD / KB)
#define A_BLK (128 * 128)
#define B_BLK (64 * 128)
#define PROD 256
#define THREADS (PROD + 32 + 128)   // producers | MMA warp | 4 epilogue warps

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
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
// arrive on the barrier at the same offset in the LEADER CTA (rank 0) of the cluster, cluster-scope release
__device__ __forceinline__ void arrive_leader(uint64_t* b) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(b)), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void bounded_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 2000000000ll) {
      printf("probe: wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
#define SPLIT(x, h, l) { h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); l = x - h; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) pipe(const float* X, const float* W, float* Y, int T) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  uint8_t* Wh = smem;                       // [NKB][8 KB]  this CTA's 64 weight rows
  uint8_t* Wl = Wh + NKB * B_BLK;
  uint8_t* ring = Wl + NKB * B_BLK;         // 2 stages x (head 16 KB | tail 16 KB)
  __shared__ uint64_t full[2], mma_done[2], acc_done[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  for (int idx = tid; idx < 64 * (KD / 4); idx += THREADS) {
    const int r = idx / (KD / 4), c4 = idx % (KD / 4), kb = c4 / 8, c = c4 % 8;
    const float4 v = *reinterpret_cast<const float4*>(W + (size_t)(rank * 64 + r) * KD + c4 * 4);
    float4 h, l;
    SPLIT(v.x, h.x, l.x) SPLIT(v.y, h.y, l.y) SPLIT(v.z, h.z, l.z) SPLIT(v.w, h.w, l.w)
    const uint32_t off = kb * B_BLK + sw128_off(r, c);
    *reinterpret_cast<float4*>(Wh + off) = h;
    *reinterpret_cast<float4*>(Wl + off) = l;
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 2 * (PROD / 32));   // producer warps of BOTH CTAs (only the leader's copy is used)
      mbar_init(&mma_done[i], 1);
      mbar_init(&acc_done[i], 1);
      mbar_init(&acc_free[i], 2 * 4);         // epilogue warps of both CTAs (leader's copy)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PROD / 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp < PROD / 32) {
    // ------------------------------------------------------------------------------------------------ producers
    const int prow = tid >> 3, c4 = tid & 7;
    unsigned cnt = 0;
    for (int t = cluster; t < T; t += nclusters) {
      const float* xb = X + ((size_t)t * 256 + rank * 128) * KD;
      for (int kb = 0; kb < NKB; ++kb, ++cnt) {
        const int stage = cnt & 1;
        if (cnt >= 2) bounded_wait(&mma_done[stage], ((cnt >> 1) - 1) & 1);
        uint8_t* sh = ring + stage * 2 * A_BLK;
        uint8_t* sl = sh + A_BLK;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int row = prow + 32 * q;
          const float4 v = *reinterpret_cast<const float4*>(xb + (size_t)row * KD + kb * KB + c4 * 4);
          float4 h, l;
          SPLIT(v.x, h.x, l.x) SPLIT(v.y, h.y, l.y) SPLIT(v.z, h.z, l.z) SPLIT(v.w, h.w, l.w)
          const uint32_t off = sw128_off(row, c4);
          *reinterpret_cast<float4*>(sh + off) = h;
          *reinterpret_cast<float4*>(sl + off) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) arrive_leader(&full[stage]);
      }
    }
  } else if (warp == PROD / 32) {
    // --------------------------------------------------------------------------------------- MMA issuer (leader)
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      unsigned cnt = 0, ti = 0;
      for (int t = cluster; t < T; t += nclusters, ++ti) {
        const uint32_t buf = ti & 1u;
        if (ti >= 2) {
          bounded_wait(&acc_free[buf], ((ti >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tacc = tmem + buf * 128u;
        for (int kb = 0; kb < NKB; ++kb, ++cnt) {
          const int stage = cnt & 1;
          bounded_wait(&full[stage], (cnt >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = smem_u32(ring + stage * 2 * A_BLK), al = ah + A_BLK;
          const uint32_t wh = smem_u32(Wh + kb * B_BLK), wl = smem_u32(Wl + kb * B_BLK);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t o = j * 32;
#define MMA2(DA, DB, ACC)                                                                                      \
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                   \
                         "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"                      \
                         ::"r"(tacc), "l"(make_desc(DA)), "l"(make_desc(DB)), "r"(idesc), "r"(ACC) : "memory")
            MMA2(ah + o, wh + o, (kb | j) ? 1u : 0u);
            MMA2(ah + o, wl + o, 1u);
            MMA2(al + o, wh + o, 1u);
          }
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                       ::"r"(smem_u32(&mma_done[stage])), "h"((uint16_t)3) : "memory");
          if (kb == NKB - 1)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&acc_done[buf])), "h"((uint16_t)3) : "memory");
        }
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------------- epilogue
    const int q = warp & 3;   // TMEM lane quarter: a warp may only touch lanes [32 (warp % 4), +32)
    unsigned ti = 0;
    for (int t = cluster; t < T; t += nclusters, ++ti) {
      const uint32_t buf = ti & 1u;
      bounded_wait(&acc_done[buf], (ti >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* yrow = Y + ((size_t)t * 256 + rank * 128 + q * 32 + lane) * N;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + buf * 128u + ((uint32_t)(q * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                       "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                       "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(yrow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) arrive_leader(&acc_free[buf]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == PROD / 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main() {
  const int T = 37;   // pair tiles of 256 rows: odd on purpose (clusters get different tile counts)
  const size_t R = (size_t)T * 256;
  std::vector<float> X(R * KD), W((size_t)N * KD), Y(R * N, -1.f);
  srand(3);
  for (auto& x : X) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : W) x = (float)rand() / RAND_MAX * 2 - 1;
  float *dX, *dW, *dY;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dY, Y.size() * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dY, 0xff, Y.size() * 4);
  const size_t smem = (size_t)2 * NKB * B_BLK + 4 * A_BLK;
  cudaFuncSetAttribute(pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int clusters : {1, 4}) {
    cudaMemset(dY, 0xff, Y.size() * 4);
    pipe<<<2 * clusters, THREADS, smem>>>(dX, dW, dY, T);
    cudaError_t e = cudaDeviceSynchronize();
    printf("2cta pipeline, %d cluster(s), %d pair tiles: %s\n", clusters, T, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < R; i += 7)   // every 7th row: all tiles, both CTAs, all lane quarters
      for (int j = 0; j < N; ++j) {
        double r = 0;
        for (int k = 0; k < KD; ++k) r += (double)X[i * KD + k] * (double)W[(size_t)j * KD + k];
        maxerr = fmax(maxerr, fabs(r - Y[i * N + j]));
        maxref = fmax(maxref, fabs(r));
      }
    printf("   max|ref| %.4f  max err %.3e  rel %.3e\n", maxref, maxerr, maxerr / maxref);
  }
  cudaFree(dX); cudaFree(dY);
  // ---- throughput at the phi size of cfg 4 (2 x 575 454 rows ~ 4496 pair tiles), one cluster per SM pair
  {
    const int Tbig = 4496;
    const size_t Rb = (size_t)Tbig * 256;
    float *bX, *bY;
    cudaMalloc(&bX, Rb * KD * 4); cudaMalloc(&bY, Rb * N * 4);
    cudaMemset(bX, 0, Rb * KD * 4);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int clusters = sms / 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) pipe<<<2 * clusters, THREADS, smem>>>(bX, dW, bY, Tbig);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) pipe<<<2 * clusters, THREADS, smem>>>(bX, dW, bY, Tbig);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / 5;
    printf("throughput: %d clusters, %zu rows x 128 -> 128 (3xTF32, no prologue / statistics, direct stores): %s, %.1f us per launch, "
           "%.0f GB/s of x+y traffic\n", clusters, Rb, cudaGetErrorString(e), us, 2.0 * Rb * 128 * 4 / (us * 1e-6) / 1e9);
    cudaFree(bX); cudaFree(bY);
  }
  cudaFree(dW);
  return 0;
}

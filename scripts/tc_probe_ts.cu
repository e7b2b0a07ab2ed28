// Standalone probe: one CTA, D[128 x 128] = A[128 x K] * B[128 x K]^T with tcgen05.mma kind::tf32 (3x split) where the
// A operand lives in TENSOR MEMORY (lane = row, one 32-bit column per K element; written with tcgen05.st) and B in shared
// memory (canonical K-major SWIZZLE_128B).  Validates the operand layout linear_tc_ws.cu relies on.
// split 1: rounded heads, split 2: raw fp32 words as heads (truncation by the tensor core).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define M 128
#define N 128
#define KB 32           // floats per K-block (128 B rows)
#define TILE_BYTES (128 * 128)  // one [128 rows x 32 floats] block

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                    // LBO (ignored for swizzled K-major) = 1
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;  // SBO = 1024 B between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

// element (row r, float4 chunk c of the 32-float K-block) -> byte offset inside the [128 x 32] block
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(256, 1) probe(const float* A, const float* B, float* D, int K, int split) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = K / KB;
  uint8_t* b_hi = smem;                       // [nkb][16 KB]
  uint8_t* b_lo = b_hi + nkb * TILE_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int idx = tid; idx < 128 * (K / 4); idx += blockDim.x) {
    const int r = idx / (K / 4), c4 = idx % (K / 4);
    const int kb = c4 / 8, c = c4 % 8;
    float4 vb = *reinterpret_cast<const float4*>(B + (size_t)r * K + c4 * 4);
    float4 bh, bl;
    #define SPLIT(x, h, l) { uint32_t u = __float_as_uint(x) & 0xFFFFE000u; h = __uint_as_float(u); l = x - h; }
    SPLIT(vb.x, bh.x, bl.x) SPLIT(vb.y, bh.y, bl.y) SPLIT(vb.z, bh.z, bl.z) SPLIT(vb.w, bh.w, bl.w)
    const uint32_t off = kb * TILE_BYTES + sw128_off(r, c);
    // split 2: heads are the RAW fp32 words, tails = x - trunc_tf32(x): correct only if the tensor core truncates
    *reinterpret_cast<float4*>(b_hi + off) = (split == 1) ? bh : vb;
    *reinterpret_cast<float4*>(b_lo + off) = bl;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  {   // A into tensor memory: warps 0..3 heads (columns 256 + k), warps 4..7 tails (columns 384 + k); lane quarter = warp & 3
    const int q = warp & 3, is_tail = warp >> 2, row = q * 32 + lane;
    for (int c = 0; c < K / 32; ++c) {
      uint32_t v[32];
      for (int j = 0; j < 32; ++j) {
        const float x = A[(size_t)row * K + c * 32 + j];
        float h, l;
        SPLIT(x, h, l)
        v[j] = __float_as_uint(is_tail ? l : (split == 1 ? h : x));
      }
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (is_tail ? 384 : 256) + c * 32;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                   "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                   ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                     "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
                     "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
                     "r"(v[30]), "r"(v[31]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    int first = 1;
    const int npass = split ? 3 : 1;
    for (int p = 0; p < npass; ++p) {
      const uint32_t pa = tmem + ((p == 2) ? 384 : 256);   // passes: hi*hi, hi*lo, lo*hi
      const uint8_t* pb = (p == 1) ? b_lo : b_hi;
      for (int kb = 0; kb < nkb; ++kb)
        for (int j = 0; j < 4; ++j) {
          const uint32_t ta = pa + kb * 32 + j * 8;
          const uint64_t db = make_desc(smem_u32(pb + kb * TILE_BYTES) + j * 32);
          const uint32_t acc = first ? 0u : 1u;
          first = 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
      if (clock64() - t0 > 1000000000ll) { if (tid == 0) printf("probe: wait timed out\n"); __trap(); }
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    const int row = warp * 32 + lane;
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  for (int K : {32, 128}) {
    for (int split = 1; split < 3; ++split) {
      std::vector<float> A(M * K), B(N * K), D(M * N, -1.f);
      srand(1);
      for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
      for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
      float *dA, *dB, *dD;
      cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
      cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
      cudaMemset(dD, 0xff, D.size() * 4);
      size_t smem = (size_t)2 * (K / KB) * TILE_BYTES + 1024;
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      probe<<<1, 256, smem>>>(dA, dB, dD, K, split);
      cudaError_t e = cudaDeviceSynchronize();
      printf("K=%d split=%d: %s\n", K, split, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int i = 0; i < M; ++i)
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

// Issue rate of the warp-level tensor-core instructions on a B200 SM (the "legacy" mma.sync path, SASS HMMA), next to FFMA:
// W warps per SM each issue L instructions over 8 independent accumulators; cycles per instruction per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/mma_sync_rate scripts/mma_sync_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void rate_kernel(float* out, long long* cyc, int L) {
  float d[8][4];
  for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 11, b0 = 13, b1 = threadIdx.x ^ 5;
  float f[32];
  for (int i = 0; i < 32; ++i) f[i] = (float)i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < L / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 2)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(b0));
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) f[i * 4 + j] = fmaf(f[i * 4 + j], 1.0001f, 0.5f);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  for (int i = 0; i < 32; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int L = 8192;
  const char* names[4] = {"mma.sync m16n8k8 tf32 ", "mma.sync m16n8k16 bf16", "mma.sync m16n8k4 tf32 ", "4 x FFMA (reg form)   "};
  for (int kind = 0; kind < 4; ++kind) {
    for (int warps : {4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (kind == 0) rate_kernel<0><<<148, warps * 32>>>(out, cyc, L);
        if (kind == 1) rate_kernel<1><<<148, warps * 32>>>(out, cyc, L);
        if (kind == 2) rate_kernel<2><<<148, warps * 32>>>(out, cyc, L);
        if (kind == 3) rate_kernel<3><<<148, warps * 32>>>(out, cyc, L);
      }
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)h[i] / 148;
      const double per_smsp = avg / ((double)L * warps / 4.0);   // cycles per instruction per sub-partition
      const double macs = kind == 0 ? 1024 : kind == 1 ? 2048 : kind == 2 ? 512 : 128;
      printf("%s  %2d warps/SM: %6.2f cycles per instruction per sub-partition  = %7.1f MAC/clk/SM\n", names[kind], warps,
             per_smsp, 4.0 * macs / per_smsp);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

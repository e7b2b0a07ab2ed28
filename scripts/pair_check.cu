// Standalone check of the experimental Linear kernels THROUGH THE C ABI, no torch.  usage: pair_check [mode]
//   mode 2 = CTA pair (csrc/linear_tc_pair.cu, default), 3 = TMA-fed, 4 = TMA-fed with raw heads (csrc/linear_tc_tma.cu),
//   5 / 6 = weight resident in tensor memory (csrc/linear_tc_ws.cu; 6 = raw heads).
//   For a list of shapes run sb_linear_fwd with the validated single-CTA tcgen05 kernel (sb_set_tensor_cores(1)) and with
//   the kernel under test (sb_set_tensor_cores(mode)) on the same device buffers and compare outputs and BatchNorm
//   statistics (modes 2 and 3: same MMAs in the same order on the same split operands -> must be bit-identical; mode 4
//   truncates the heads instead of rounding them -> fp64 bound only), spot-check rows against an fp64 host product,
//   then time both at the phi size of cfg 4.
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/build/pair_check scripts/pair_check.cu \
//             -Lsignnet_basisnet_b200/lib -lsignnet_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../signnet_basisnet_b200/lib'
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../include/signnet_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void fill(float* p, size_t n, uint32_t seed, float scale, float shift) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    p[i] = ((float)(h >> 8) / 16777216.f * 2.f - 1.f) * scale + shift;
  }
}

struct Case { long long R; int G, K, N, pro, relu, stats, bias, wt; };

static int run(const Case& c, int mode, const float* x, const float* w, const float* b, const float* pa, const float* pc,
               float* y, double* st) {
  sb_set_tensor_cores(mode);
  CK(cudaMemsetAsync(st, 0, sizeof(double) * 2 * 2 * 128));
  // wt: the weight is stored [K, N] (input-gradient orientation): row stride 1, column stride N
  const long long w_rs = c.wt ? 1 : c.K, w_cs = c.wt ? c.N : 1;
  return sb_linear_fwd(x, c.K, w, w_rs, w_cs, c.bias ? b : nullptr, y, c.N, c.R, c.G, c.K, c.N, c.pro, pa, pc, c.relu,
                       c.stats ? st : nullptr, 0, nullptr);
}

int main(int argc, char** argv) {
  const int mode_ut = (argc > 1) ? atoi(argv[1]) : 2;
  const bool exact = mode_ut == 2 || mode_ut == 3;   // 4, 6: truncated heads; 5, 6: transposed MMA (order inside the tensor core unknown)
  printf("kernel under test: sb_set_tensor_cores(%d)\n", mode_ut);
  const Case cases[] = {
      {8200, 1, 32, 32, 0, 0, 1, 0, 0},       // 65 tiles: odd -> the peer's last tile is dead
      {33333, 2, 64, 96, 1, 1, 1, 1, 0},      // ragged last tile per group, prologue affine, relu, bias, statistics
      {70001, 2, 96, 64, 2, 1, 1, 1, 1},      // transposed weight, prologue affine + relu
      {12345, 1, 128, 32, 0, 0, 0, 1, 0},
      {20000, 2, 128, 128, 2, 0, 1, 1, 0},
      {575454, 2, 128, 128, 2, 0, 1, 1, 0},   // phi size of cfg 4: second Linear of MaskedMLP (timed)
      {575454, 2, 128, 128, 0, 0, 0, 0, 1},   // phi size: input gradient (timed)
  };
  const size_t maxel = (size_t)2 * 575454 * 128;
  float *x, *y1, *y2, *w, *b, *pa, *pc;
  double* st;
  CK(cudaMalloc(&x, maxel * 4)); CK(cudaMalloc(&y1, maxel * 4)); CK(cudaMalloc(&y2, maxel * 4));
  CK(cudaMalloc(&w, 128 * 128 * 4)); CK(cudaMalloc(&b, 128 * 4)); CK(cudaMalloc(&pa, 256 * 4)); CK(cudaMalloc(&pc, 256 * 4));
  CK(cudaMalloc(&st, sizeof(double) * 2 * 2 * 128));
  fill<<<1184, 256>>>(x, maxel, 1u, 1.f, 0.f);
  fill<<<64, 256>>>(w, 128 * 128, 2u, 0.2f, 0.f);
  fill<<<1, 128>>>(b, 128, 3u, 0.5f, 0.f);
  fill<<<1, 256>>>(pa, 256, 4u, 0.5f, 1.f);
  fill<<<1, 256>>>(pc, 256, 5u, 0.3f, 0.f);
  CK(cudaDeviceSynchronize());
  std::vector<float> hw(128 * 128), hb(128), hpa(256), hpc(256);
  CK(cudaMemcpy(hw.data(), w, hw.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), b, 512, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hpa.data(), pa, 1024, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hpc.data(), pc, 1024, cudaMemcpyDeviceToHost));

  if (argc > 2) {   // "pair_check <mode> <case index>": launch the kernel under test three times on one case (for ncu)
    const Case& c = cases[atoi(argv[2])];
    for (int i = 0; i < 3; ++i) {
      const int rc = run(c, mode_ut, x, w, b, pa, pc, y2, st);
      if (rc || sb_last_linear_kernel() != mode_ut) { printf("rc %d kernel %d %s\n", rc, sb_last_linear_kernel(), sb_last_error()); return 1; }
    }
    CK(cudaDeviceSynchronize());
    printf("ran case %s with kernel %d\n", argv[2], mode_ut);
    return 0;
  }
  int bad = 0;
  for (const Case& c : cases) {
    const size_t nel = (size_t)c.G * c.R * c.N;
    double s1[512], s2[512];
    CK(cudaMemset(y1, 0xff, nel * 4)); CK(cudaMemset(y2, 0xff, nel * 4));
    int rc = run(c, 1, x, w, b, pa, pc, y1, st);
    if (rc) { printf("single-CTA: rc %d %s\n", rc, sb_last_error()); return 1; }
    CK(cudaMemcpy(s1, st, sizeof(s1), cudaMemcpyDeviceToHost));
    rc = run(c, mode_ut, x, w, b, pa, pc, y2, st);
    if (rc) { printf("under test: rc %d %s\n", rc, sb_last_error()); return 1; }
    if (sb_last_linear_kernel() != mode_ut) { printf("the dispatcher fell back to kernel %d\n", sb_last_linear_kernel()); return 1; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel under test failed: %s\n", cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(s2, st, sizeof(s2), cudaMemcpyDeviceToHost));
    std::vector<float> h1(nel), h2(nel);
    CK(cudaMemcpy(h1.data(), y1, nel * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), y2, nel * 4, cudaMemcpyDeviceToHost));
    size_t ndiff = 0; double maxd = 0;
    for (size_t i = 0; i < nel; ++i)
      if (memcmp(&h1[i], &h2[i], 4)) { ++ndiff; maxd = fmax(maxd, fabs((double)h1[i] - (double)h2[i])); }
    double sd = 0;
    if (c.stats) for (int i = 0; i < c.G * 2 * c.N; ++i) sd = fmax(sd, fabs(s1[i] - s2[i]) / (fabs(s1[i]) + 1e-30));
    // fp64 spot check of the pair kernel: 64 rows spread over the row space
    std::vector<float> hx(c.K);
    double maxerr = 0, maxref = 0;
    for (int s = 0; s < 64; ++s) {
      const long long row = (long long)((double)s / 63.0 * (double)(c.G * c.R - 1));
      const int g = (int)(row / c.R);
      CK(cudaMemcpy(hx.data(), x + row * c.K, c.K * 4, cudaMemcpyDeviceToHost));
      for (int n = 0; n < c.N; ++n) {
        double acc = c.bias ? hb[n] : 0.0;
        for (int k = 0; k < c.K; ++k) {
          float v = hx[k];
          if (c.pro) { v = fmaf(hpa[g * c.K + k], v, hpc[g * c.K + k]); if (c.pro == 2) v = fmaxf(v, 0.f); }
          acc += (double)v * (double)(c.wt ? hw[(size_t)k * c.N + n] : hw[(size_t)n * c.K + k]);
        }
        if (c.relu) acc = fmax(acc, 0.0);
        maxerr = fmax(maxerr, fabs(acc - (double)h2[row * c.N + n]));
        maxref = fmax(maxref, fabs(acc));
      }
    }
    const bool ok = (exact ? (ndiff == 0 && sd < 1e-12) : (maxd <= 2e-5 * fmax(maxref, 1.0) && sd < 1e-5)) &&
                    maxerr <= 2e-5 * fmax(maxref, 1.0);
    if (!ok) ++bad;
    printf("R=%lld G=%d K=%d N=%d pro=%d relu=%d stats=%d bias=%d wt=%d : %zu/%zu elements differ (max %.3e), stats rel diff %.2e, "
           "vs fp64 err %.3e (max|ref| %.3f)  %s\n", c.R, c.G, c.K, c.N, c.pro, c.relu, c.stats, c.bias, c.wt, ndiff, nel, maxd, sd,
           maxerr, maxref, ok ? "OK" : "MISMATCH");
    if (c.R >= 500000) {
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      for (int mi = 0; mi < 2; ++mi) {
        const int mode = mi ? mode_ut : 1;
        for (int i = 0; i < 3; ++i) run(c, mode, x, w, b, pa, pc, y2, st);
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 10; ++i) run(c, mode, x, w, b, pa, pc, y2, st);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 100.0;
        printf("   %s: %.1f us per launch (incl. the 4 KB statistics memset), %.0f GB/s of x + y traffic\n",
               mode == 1 ? "single-CTA" : "under test", us, (double)c.G * c.R * (c.K + c.N) * 4 / (us * 1e-6) / 1e9);
      }
    }
  }
  // ---- weight gradient: the TMA-fed variant (modes >= 3, N == K == 128) against the default tcgen05 wgrad kernel
  if (mode_ut >= 3) {
    const bool wexact = mode_ut == 3 || mode_ut == 5;   // rounded heads: same operands, same MMA order
    float *ws, *dw1, *dw2, *db1, *db2;
    CK(cudaMalloc(&ws, sb_linear_wgrad_workspace_floats() * 4));
    CK(cudaMalloc(&dw1, 128 * 128 * 4)); CK(cudaMalloc(&dw2, 128 * 128 * 4)); CK(cudaMalloc(&db1, 512)); CK(cudaMalloc(&db2, 512));
    struct WCase { long long R; int G, pro; };
    const WCase wcases[] = {{20011, 2, 2}, {20011, 1, 0}, {575454, 2, 2}, {575454, 2, 0}};
    for (const WCase& c : wcases) {
      // g = y1 region (refilled), x = x
      fill<<<1184, 256>>>(y1, (size_t)c.G * c.R * 128, 7u, 1.f, 0.f);
      std::vector<float> h1(128 * 128), h2(128 * 128), b1(128), b2(128);
      sb_set_tensor_cores(1);
      int rc = sb_linear_wgrad(y1, 128, x, 128, c.R, c.G, 128, 128, c.pro, pa, pc, dw1, 128, 1, db1, 0, ws, nullptr);
      if (rc || sb_last_wgrad_kernel() != 1) { printf("wgrad default: rc %d kernel %d %s\n", rc, sb_last_wgrad_kernel(), sb_last_error()); return 1; }
      sb_set_tensor_cores(mode_ut);
      rc = sb_linear_wgrad(y1, 128, x, 128, c.R, c.G, 128, 128, c.pro, pa, pc, dw2, 128, 1, db2, 0, ws, nullptr);
      if (rc || sb_last_wgrad_kernel() != mode_ut) { printf("wgrad under test: rc %d kernel %d %s\n", rc, sb_last_wgrad_kernel(), sb_last_error()); return 1; }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("wgrad kernel under test failed: %s\n", cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(h1.data(), dw1, 65536, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), dw2, 65536, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(b1.data(), db1, 512, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b2.data(), db2, 512, cudaMemcpyDeviceToHost));
      size_t ndiff = 0; double maxd = 0, maxref = 0, maxdb = 0;
      for (int i = 0; i < 128 * 128; ++i) {
        if (memcmp(&h1[i], &h2[i], 4)) ++ndiff;
        maxd = fmax(maxd, fabs((double)h1[i] - (double)h2[i]));
        maxref = fmax(maxref, fabs((double)h1[i]));
      }
      for (int i = 0; i < 128; ++i) maxdb = fmax(maxdb, fabs((double)b1[i] - (double)b2[i]));
      double maxerr = -1;
      if (c.R < 100000) {   // fp64 host product at the small size
        std::vector<float> hg((size_t)c.G * c.R * 128), hx((size_t)c.G * c.R * 128);
        CK(cudaMemcpy(hg.data(), y1, hg.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hx.data(), x, hx.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<double> ref(128 * 128, 0.0);
        std::vector<float> fx(128);
        for (long long r = 0; r < c.G * c.R; ++r) {
          const int g = (int)(r / c.R);
          for (int k = 0; k < 128; ++k) {
            float v = hx[r * 128 + k];
            if (c.pro) { v = fmaf(hpa[g * 128 + k], v, hpc[g * 128 + k]); if (c.pro == 2) v = fmaxf(v, 0.f); }
            fx[k] = v;
          }
          for (int n = 0; n < 128; ++n) {
            const double gv = hg[r * 128 + n];
            double* rr = &ref[n * 128];
            for (int k = 0; k < 128; ++k) rr[k] += gv * (double)fx[k];
          }
        }
        maxerr = 0;
        for (int i = 0; i < 128 * 128; ++i) maxerr = fmax(maxerr, fabs(ref[i] - (double)h2[i]));
      }
      const bool ok = (wexact ? (ndiff == 0 && maxdb == 0) : (maxd <= 2e-5 * fmax(maxref, 1.0))) &&
                      (maxerr < 0 || maxerr <= 2e-5 * fmax(maxref, 1.0));
      if (!ok) ++bad;
      printf("wgrad R=%lld G=%d pro=%d : %zu/16384 elements differ (max %.3e, max|dw| %.3f), max db diff %.3e, vs fp64 err %.3e  %s\n",
             c.R, c.G, c.pro, ndiff, maxd, maxref, maxdb, maxerr, ok ? "OK" : "MISMATCH");
      if (c.R >= 500000) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int mi = 0; mi < 2; ++mi) {
          sb_set_tensor_cores(mi ? mode_ut : 1);
          for (int i = 0; i < 3; ++i) sb_linear_wgrad(y1, 128, x, 128, c.R, c.G, 128, 128, c.pro, pa, pc, dw2, 128, 1, db2, 0, ws, nullptr);
          CK(cudaEventRecord(e0));
          for (int i = 0; i < 10; ++i) sb_linear_wgrad(y1, 128, x, 128, c.R, c.G, 128, 128, c.pro, pa, pc, dw2, 128, 1, db2, 0, ws, nullptr);
          CK(cudaEventRecord(e1));
          CK(cudaDeviceSynchronize());
          float ms = 0;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          printf("   %s: %.1f us per call (wgrad + partial reduction), %.0f GB/s of g + x traffic\n", mi ? "under test" : "default   ",
                 ms * 100.0, (double)c.G * c.R * 256 * 4 / (ms * 100.0 * 1e-6) / 1e9);
        }
      }
    }
  }
  printf(bad ? "pair_check: %d case(s) MISMATCH\n" : "pair_check: all cases OK\n", bad);
  return bad ? 1 : 0;
}

// K9 — the five 2->1 equivariant contractions of BasisNet's IGN phi (LearningFilters/ign.py:344-374
// contractions_2_to_1, normalization='inf') on eigenspace projectors P_e = V_e V_e^T:
//     ops[e, i, :] = { P_ii, tr(P)/n, sum_j P_ij / n, sum_j P_ji / n, sum_ij P_ij / n^2 }
// The reference materialises every projector ([b,1,n,n] fp32: 4 b n^2 bytes, 2.1 GB for the 32x32 grid) and reduces it
// with five torch ops.  P is a rank-mult symmetric product, so all five follow from the eigenvector block V_e [n, mult]:
//     P_ii = |V_e[i,:]|^2      sum_j P_ij = V_e[i,:] . s_e   (s_e = V_e^T 1)      tr = sum_i P_ii      total = |s_e|^2
// i.e. 4 n mult bytes per eigenspace instead of 4 n^2 - the projector never exists.  A second kernel takes materialised
// projectors for callers that already hold them (the reference's input format).
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define IGN_THREADS 256
#define IGN_MAXMULT 64

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];   // fixed order
  return t;
}

// one CTA per eigenspace e: columns [col0[e], col0[e] + mult) of V [n, ldv]
__global__ void __launch_bounds__(IGN_THREADS) ign_ops_factors_kernel(const float* __restrict__ V, long long ldv, int n,
                                                                      const int32_t* __restrict__ col0, int mult,
                                                                      float* __restrict__ ops, int ldo) {
  __shared__ double sh[IGN_THREADS / 32];
  __shared__ float s_col[IGN_MAXMULT];
  __shared__ float s_tr, s_tot;
  const int e = blockIdx.x;
  const float* Ve = V + col0[e];
  // pass 1: column sums s_m and the trace
  double tr = 0.0;
  for (int m = 0; m < mult; ++m) {
    double cs = 0.0;
    for (int i = threadIdx.x; i < n; i += IGN_THREADS) {
      const float v = __ldg(Ve + (long long)i * ldv + m);
      cs += (double)v;
      tr += (double)v * (double)v;
    }
    const double t = block_sum_d(cs, sh);
    if (threadIdx.x == 0) s_col[m] = (float)t;
  }
  const double trs = block_sum_d(tr, sh);
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int m = 0; m < mult; ++m) tot += (double)s_col[m] * (double)s_col[m];
    s_tr = (float)(trs / (double)n);
    s_tot = (float)(tot / ((double)n * (double)n));
  }
  __syncthreads();
  // pass 2: per-row diagonal and row sum (V_e stays in L1/L2: n * mult floats)
  const float inv_n = 1.0f / (float)n;
  for (int i = threadIdx.x; i < n; i += IGN_THREADS) {
    float d = 0.f, r = 0.f;
    for (int m = 0; m < mult; ++m) {
      const float v = __ldg(Ve + (long long)i * ldv + m);
      d = fmaf(v, v, d);
      r = fmaf(v, s_col[m], r);
    }
    float* o = ops + ((long long)e * n + i) * ldo;
    o[0] = d;
    o[1] = s_tr;
    o[2] = r * inv_n;
    o[3] = r * inv_n;   // P symmetric: column sums = row sums
    o[4] = s_tot;
    for (int c = 5; c < ldo; ++c) o[c] = 0.f;
  }
}

// materialised projectors P [b, n, n]: a warp per row (row sums, diagonal), then a thread per column (column sums)
__global__ void __launch_bounds__(IGN_THREADS) ign_ops_rows_kernel(const float* __restrict__ P, int n,
                                                                   float* __restrict__ ops, int ldo,
                                                                   double* __restrict__ acc /*[b][2]: trace, total*/) {
  const int e = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (IGN_THREADS / 32) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (IGN_THREADS / 32);
  const float* Pe = P + (long long)e * n * n;
  double tr = 0.0, tot = 0.0;
  for (int i = warp; i < n; i += nwarps) {
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += __ldg(Pe + (long long)i * n + j);
    s = warp_sum(s);
    if (lane == 0) {
      float* o = ops + ((long long)e * n + i) * ldo;
      const float d = __ldg(Pe + (long long)i * n + i);
      o[0] = d;
      o[2] = s / (float)n;
      tr += (double)d;
      tot += (double)s;
    }
  }
  if (lane == 0 && (tr != 0.0 || tot != 0.0)) {
    atomicAdd(acc + 2 * e, tr);
    atomicAdd(acc + 2 * e + 1, tot);
  }
}
__global__ void __launch_bounds__(IGN_THREADS) ign_ops_cols_kernel(const float* __restrict__ P, int n,
                                                                   float* __restrict__ ops, int ldo,
                                                                   const double* __restrict__ acc) {
  const int e = blockIdx.y;
  const int j = blockIdx.x * IGN_THREADS + threadIdx.x;
  if (j >= n) return;
  const float* Pe = P + (long long)e * n * n;
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += __ldg(Pe + (long long)i * n + j);
  float* o = ops + ((long long)e * n + j) * ldo;
  o[1] = (float)(acc[2 * e] / (double)n);
  o[3] = s / (float)n;
  o[4] = (float)(acc[2 * e + 1] / ((double)n * (double)n));
  for (int c = 5; c < ldo; ++c) o[c] = 0.f;
}

extern "C" int sb_ign2to1_ops_factors(const float* V, int64_t ldv, int32_t n, const int32_t* col0, int32_t b,
                                      int32_t mult, float* ops, int32_t ldo, void* stream) {
  SB_CHECK_ARG(n >= 1 && b >= 0 && mult >= 1 && mult <= IGN_MAXMULT && ldo >= 5 && ldv >= mult,
               "sb_ign2to1_ops_factors: bad sizes n=%d b=%d mult=%d ldo=%d", n, b, mult, ldo);
  if (b == 0) return SB_OK;
  ign_ops_factors_kernel<<<b, IGN_THREADS, 0, (cudaStream_t)stream>>>(V, ldv, n, col0, mult, ops, ldo);
  SB_CHECK_LAUNCH("sb_ign2to1_ops_factors");
  return SB_OK;
}

extern "C" int sb_ign2to1_ops_projectors(const float* P, int32_t n, int32_t b, float* ops, int32_t ldo,
                                         double* workspace /*[2 b], zeroed here*/, void* stream) {
  SB_CHECK_ARG(n >= 1 && b >= 0 && ldo >= 5 && workspace, "sb_ign2to1_ops_projectors: bad arguments");
  if (b == 0) return SB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * b, st));
  int gx = (n + (IGN_THREADS / 32) * 4 - 1) / ((IGN_THREADS / 32) * 4);
  if (gx < 1) gx = 1;
  ign_ops_rows_kernel<<<dim3(gx, b), IGN_THREADS, 0, st>>>(P, n, ops, ldo, workspace);
  SB_CHECK_LAUNCH("sb_ign2to1_ops_projectors(rows)");
  ign_ops_cols_kernel<<<dim3((n + IGN_THREADS - 1) / IGN_THREADS, b), IGN_THREADS, 0, st>>>(P, n, ops, ldo, workspace);
  SB_CHECK_LAUNCH("sb_ign2to1_ops_projectors(cols)");
  return SB_OK;
}

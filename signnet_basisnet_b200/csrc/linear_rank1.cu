// Degenerate contractions of the first phi layer (d_in = 1: the eigenvector entry itself,
// Alchemy/sign_net/sign_net.py:18-21 builds MaskedGINConv(1, n_hid)): K = 1 or N = 1 turns Linear into an outer
// product / a row dot product / a column-weighted sum.  They carry no tensor-core work at all - each is one streaming
// pass over a [rows, N] activation (HBM bound) - so they get their own kernels instead of the tiled contraction:
//   rank1_fwd_kernel    y[r, n]  = f(x[r]) * w[n] + b[n]            (+ ReLU, + fp64 column statistics)   writes T
//   rowdot_fwd_kernel   y[r]     = sum_k f(x[r, k]) * w[k] + b       (+ ReLU)                              reads  T
//   rank1_wgrad_kernel  dw[n]    = sum_r g[r, n] * f(x[r]),  db[n] = sum_r g[r, n]                         reads  T
// Same argument contract as sb_linear_fwd / sb_linear_wgrad (include/signnet_b200.h); dispatched from linear.cu.
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define R1_THREADS 256
#define R1_MAXG 2

struct R1Args {
  const float* x;      // [G*R] (ldx) scalar per row            | rowdot: [G*R, ldx]
  long long ldx;
  const float* w;      // K=1: w[n * w_stride]                  | rowdot: w[k * w_stride]
  long long w_stride;
  const float* bias;
  float* y;            // [G*R, ldy]                            | rowdot: [G*R] (ldy)
  long long ldy;
  long long R;
  int G, N, ycols;     // rowdot: N = K (columns of x)
  int pro;
  const float* pa;     // [G, K]
  const float* pc;
  int relu;
  double* stats;       // [G, 2, N] or null
};

__device__ __forceinline__ float r1_pro(float v, int pro, float pa, float pc) {
  if (pro) {
    v = fmaf(pa, v, pc);
    if (pro == 2) v = fmaxf(v, 0.f);
  }
  return v;
}

// thread -> float4 column group (tid % cg) of rows (tid / cg) + k * rows_per_block; cg = ycols / 4 <= 32
__global__ void __launch_bounds__(R1_THREADS) rank1_fwd_kernel(const R1Args a) {
  __shared__ double red[R1_THREADS / 32][2][128];
  const int cg = a.ycols >> 2;
  const int rpb = R1_THREADS / cg;                 // rows per block step
  const int c4 = threadIdx.x % cg, rl = threadIdx.x / cg;
  const bool active = rl < rpb;
  float wv[4], bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = c4 * 4 + j;
    wv[j] = (n < a.N) ? __ldg(a.w + n * a.w_stride) : 0.f;
    bv[j] = (a.bias && n < a.N) ? __ldg(a.bias + n) : 0.f;
  }
  for (int g = 0; g < a.G; ++g) {
    const float pa = a.pro ? __ldg(a.pa + g) : 1.f, pc = a.pro ? __ldg(a.pc + g) : 0.f;
    double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
    if (active) {
      for (long long r = (long long)blockIdx.x * rpb + rl; r < a.R; r += (long long)gridDim.x * rpb) {
        const long long row = (long long)g * a.R + r;
        const float xv = r1_pro(__ldg(a.x + row * a.ldx), a.pro, pa, pc);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = __fadd_rn(__fmul_rn(xv, wv[j]), bv[j]);
          if (a.relu) t = fmaxf(t, 0.f);
          o[j] = (c4 * 4 + j < a.N) ? t : 0.f;
          if (a.stats) {
            s[j] += (double)o[j];
            q[j] += (double)o[j] * (double)o[j];
          }
        }
        stg4_stream(a.y + row * a.ldy + c4 * 4, make_float4(o[0], o[1], o[2], o[3]));
      }
    }
    if (a.stats) {
      // block reduction in a fixed order, then one fp64 atomic per (block, column)
      __syncthreads();
      double* rs = &red[0][0][0];
      // reuse red as [rpb rows][2][ycols] when it fits, else fall back to per-thread atomics
      if (rpb * 2 * a.ycols <= (R1_THREADS / 32) * 2 * 128) {
        if (active) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rs[(rl * 2 + 0) * a.ycols + c4 * 4 + j] = s[j];
            rs[(rl * 2 + 1) * a.ycols + c4 * 4 + j] = q[j];
          }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < 2 * a.N; idx += R1_THREADS) {
          const int which = idx / a.N, n = idx - which * a.N;
          double t = 0.0;
          for (int rr = 0; rr < rpb; ++rr) t += rs[(rr * 2 + which) * a.ycols + n];
          if (t != 0.0) atomicAdd(a.stats + ((long long)g * 2 + which) * a.N + n, t);
        }
      } else if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c4 * 4 + j < a.N) {
            atomicAdd(a.stats + ((long long)g * 2 + 0) * a.N + c4 * 4 + j, s[j]);
            atomicAdd(a.stats + ((long long)g * 2 + 1) * a.N + c4 * 4 + j, q[j]);
          }
        }
      }
      __syncthreads();
    }
  }
}

// warp per row (two rows in flight), lanes over float4 columns
__global__ void __launch_bounds__(R1_THREADS) rowdot_fwd_kernel(const R1Args a) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (R1_THREADS / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (R1_THREADS / 32);
  const int K = a.N, k4 = (K + 3) >> 2;
  const float b = a.bias ? __ldg(a.bias) : 0.f;
  const long long total = (long long)a.G * a.R;
  for (long long row = warp0; row < total; row += nwarps) {
    const int g = (row >= a.R) ? 1 : 0;             // G <= 2
    const float* xr = a.x + row * a.ldx;
    float acc = 0.f;
    for (int c = lane; c < k4; c += 32) {
      const float4 v = ldg4(xr + c * 4);
      const float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = c * 4 + j;
        if (k < K) {
          const float xv = r1_pro(t[j], a.pro, a.pro ? __ldg(a.pa + g * K + k) : 1.f, a.pro ? __ldg(a.pc + g * K + k) : 0.f);
          acc = fmaf(xv, __ldg(a.w + k * a.w_stride), acc);
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float t = acc + b;
      if (a.relu) t = fmaxf(t, 0.f);
      a.y[row * a.ldy] = t;
    }
  }
}

struct R1WgArgs {
  const float* g;      // [G*R, ldg]
  long long ldg;
  const float* x;      // [G*R] (ldx)
  long long ldx;
  long long R;
  int G, N;
  int pro;
  const float* pa;
  const float* pc;
  float* part;         // [grid][2][128]
};

__global__ void __launch_bounds__(R1_THREADS) rank1_wgrad_kernel(const R1WgArgs a) {
  __shared__ float red[R1_THREADS / 32][2][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n4 = (a.N + 3) >> 2;                    // <= 32 column groups: lane -> group
  float sw[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
  const long long total = (long long)a.G * a.R;
  const long long warp0 = (long long)blockIdx.x * (R1_THREADS / 32) + warp;
  const long long nwarps = (long long)gridDim.x * (R1_THREADS / 32);
  if (lane < n4) {
    for (long long row = warp0; row < total; row += nwarps) {
      const int g = (row >= a.R) ? 1 : 0;
      const float xv = r1_pro(__ldg(a.x + row * a.ldx), a.pro, a.pro ? __ldg(a.pa + g) : 1.f, a.pro ? __ldg(a.pc + g) : 0.f);
      const float4 v = ldg4(a.g + row * a.ldg + lane * 4);
      sw[0] = fmaf(v.x, xv, sw[0]); sw[1] = fmaf(v.y, xv, sw[1]); sw[2] = fmaf(v.z, xv, sw[2]); sw[3] = fmaf(v.w, xv, sw[3]);
      sb[0] += v.x; sb[1] += v.y; sb[2] += v.z; sb[3] += v.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[warp][0][(lane * 4 + j) & 127] = sw[j];
    red[warp][1][(lane * 4 + j) & 127] = sb[j];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * 128; idx += R1_THREADS) {
    const int which = idx >> 7, n = idx & 127;
    float t = 0.f;
    for (int w = 0; w < R1_THREADS / 32; ++w) t += red[w][which][n];
    a.part[((size_t)blockIdx.x * 2 + which) * 128 + n] = t;
  }
}

// fp64 sum of the per-block partials: a warp per output element (lanes stride the partials, fixed shuffle tree), so the
// reduction over ~1200 CTAs is 37 dependent loads deep instead of 1184
__global__ void rank1_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int N, float* __restrict__ dw,
                                          long long dw_stride, float* __restrict__ db, int accumulate) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (idx >= 2 * N) return;   // warp-uniform
  const int which = idx / N, n = idx - which * N;
  if (which == 1 && !db) return;
  double t = 0.0;
  for (int p = lane; p < nparts; p += 32) t += (double)part[((size_t)p * 2 + which) * 128 + n];
  t = warp_sum_d(t);
  if (lane == 0) {
    float* dst = which ? db + n : dw + n * dw_stride;
    *dst = accumulate ? *dst + (float)t : (float)t;
  }
}

// ---- host entry points (called from linear.cu; SB_ERR_UNSUPPORTED = use the generic path) ----------------------------
int sb_rank1_fwd_launch(const float* x, int64_t ldx, const float* w, int64_t w_rs, const float* bias, float* y,
                        int64_t ldy, int64_t R, int32_t G, int32_t N, int32_t ycols, int32_t pro, const float* pa,
                        const float* pc, int32_t relu, double* stats, cudaStream_t st) {
  if (N > 128 || G > R1_MAXG || ycols % 4 != 0 || ycols < 4 || ycols > 128 || ldy % 4 != 0 || (uintptr_t)y % 16 != 0 ||
      R * G < 4096)
    return SB_ERR_UNSUPPORTED;
  R1Args a;
  a.x = x; a.ldx = ldx; a.w = w; a.w_stride = w_rs; a.bias = bias; a.y = y; a.ldy = ldy; a.R = R; a.G = G; a.N = N;
  a.ycols = ycols; a.pro = pro; a.pa = pa; a.pc = pc; a.relu = relu; a.stats = stats;
  const int rpb = R1_THREADS / (ycols / 4);
  long long grid = sb_ceil_div(R, rpb);
  const long long cap = (long long)sb_num_sms() * 8;
  if (grid > cap) grid = cap;
  rank1_fwd_kernel<<<(unsigned)grid, R1_THREADS, 0, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_fwd(rank1)");
  return SB_OK;
}

int sb_rowdot_fwd_launch(const float* x, int64_t ldx, const float* w, int64_t w_cs, const float* bias, float* y,
                         int64_t ldy, int64_t R, int32_t G, int32_t K, int32_t pro, const float* pa, const float* pc,
                         int32_t relu, cudaStream_t st) {
  if (G > R1_MAXG || ldx % 4 != 0 || (uintptr_t)x % 16 != 0 || ldx < (K + 3) / 4 * 4 || R * G < 4096)
    return SB_ERR_UNSUPPORTED;
  R1Args a;
  a.x = x; a.ldx = ldx; a.w = w; a.w_stride = w_cs; a.bias = bias; a.y = y; a.ldy = ldy; a.R = R; a.G = G; a.N = K;
  a.ycols = 0; a.pro = pro; a.pa = pa; a.pc = pc; a.relu = relu; a.stats = nullptr;
  long long grid = sb_ceil_div(R * G, R1_THREADS / 32);
  const long long cap = (long long)sb_num_sms() * 8;
  if (grid > cap) grid = cap;
  rowdot_fwd_kernel<<<(unsigned)grid, R1_THREADS, 0, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_fwd(rowdot)");
  return SB_OK;
}

int sb_rank1_wgrad_launch(const float* gy, int64_t ldg, const float* x, int64_t ldx, int64_t R, int32_t G, int32_t N,
                          int32_t pro, const float* pa, const float* pc, float* dw, int64_t dw_rs, float* db,
                          int32_t accumulate, float* workspace, cudaStream_t st) {
  if (N > 128 || G > R1_MAXG || ldg % 4 != 0 || (uintptr_t)gy % 16 != 0 || ldg < (N + 3) / 4 * 4 || R * G < 4096)
    return SB_ERR_UNSUPPORTED;
  R1WgArgs a;
  a.g = gy; a.ldg = ldg; a.x = x; a.ldx = ldx; a.R = R; a.G = G; a.N = N; a.pro = pro; a.pa = pa; a.pc = pc;
  a.part = workspace;
  const int grid = sb_num_sms() * 8;   // 2 * 128 floats per block: well inside sb_linear_wgrad_workspace_floats()
  rank1_wgrad_kernel<<<grid, R1_THREADS, 0, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_wgrad(rank1)");
  rank1_wgrad_reduce_kernel<<<(2 * N * 32 + 127) / 128, 128, 0, st>>>(workspace, grid, N, dw, dw_rs, db, accumulate);
  SB_CHECK_LAUNCH("sb_linear_wgrad(rank1 reduce)");
  return SB_OK;
}

// K3/K4 — BatchNorm over the valid slot rows (per sign pass) and the element-wise glue of a GIN layer.
// Replaces MaskedBN's `x[mask] = bn(x[mask])` gather/scatter copies and the `x[~mask] = 0` index_put_ bookkeeping
// (Alchemy/sign_net/model_utils/masked_layers.py:13-20,59-60; sign_net.py:38-43): on the ragged slot-row layout the
// statistics are plain column sums (accumulated in fp64 by the producing Linear's epilogue, see linear.cu), the
// normalisation collapses to a per-(sign, channel) affine a*x + c that is applied in the consumer's prologue, and only
// the residual stream X_{l+1} = relu(a*Y + c) + X_l is materialised.
//
// Numerics follow nn.BatchNorm1d: biased variance for normalisation, unbiased for the running buffer, eps 1e-5,
// momentum 0.1, running statistics updated once per sign pass in +v, -v order (sign_net.py:113).
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define EW_THREADS 256

// ------------------------------------------------------------------------------------------------------- finalize
__global__ void bn_finalize_kernel(const double* __restrict__ stats, long long M, int G, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, int training, float* __restrict__ a, float* __restrict__ c,
                                   double* __restrict__ mean_rstd /*[2,G,C] fp64, for the backward*/) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  const double gm = gamma ? (double)gamma[ch] : 1.0, bt = beta ? (double)beta[ch] : 0.0;
  for (int g = 0; g < G; ++g) {
    double mean, var;
    if (training) {
      const double s = stats[((long long)g * 2 + 0) * C + ch], q = stats[((long long)g * 2 + 1) * C + ch];
      mean = s / (double)M;
      var = q / (double)M - mean * mean;
      if (var < 0.0) var = 0.0;
      if (running_mean) {
        const double unb = (M > 1) ? var * ((double)M / (double)(M - 1)) : var;
        running_mean[ch] = (float)((1.0 - (double)momentum) * (double)running_mean[ch] + (double)momentum * mean);
        running_var[ch] = (float)((1.0 - (double)momentum) * (double)running_var[ch] + (double)momentum * unb);
      }
    } else {
      mean = (double)running_mean[ch];
      var = (double)running_var[ch];
    }
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double av = gm * rstd;
    a[(long long)g * C + ch] = (float)av;
    c[(long long)g * C + ch] = (float)(bt - mean * av);
    if (mean_rstd) {
      mean_rstd[(long long)g * C + ch] = mean;
      mean_rstd[((long long)G + g) * C + ch] = rstd;
    }
  }
}

extern "C" int sb_bn_finalize(const double* stats, int64_t M, int32_t G, int32_t C, const float* gamma,
                              const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                              int32_t training, float* a, float* c, double* mean_rstd, void* stream) {
  SB_CHECK_ARG(G >= 1 && C >= 1 && a && c, "sb_bn_finalize: bad args");
  SB_CHECK_ARG(training ? (stats != nullptr && M >= 1) : (running_mean && running_var),
               "sb_bn_finalize: training needs stats and M>=1, eval needs running statistics");
  bn_finalize_kernel<<<(unsigned)sb_ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(
      stats, M, G, C, gamma, beta, running_mean, running_var, momentum, eps, training, a, c, mean_rstd);
  SB_CHECK_LAUNCH("sb_bn_finalize");
  return SB_OK;
}

// ----------------------------------------------------------------------------------- per-(group, channel) column sums
// stats[g, 0, c] += sum_r x[g, r, c] ; stats[g, 1, c] += sum_r x[g, r, c]^2     (for tensors no Linear epilogue saw)
template <int VEC>
__global__ void __launch_bounds__(EW_THREADS) col_stats_kernel(const float* __restrict__ x, long long ld, long long R,
                                                               int C, double* __restrict__ stats) {
  extern __shared__ double red[];  // [rows_per_iter][2][ldv*VEC]
  const int ldv = (int)(ld / VEC);
  const int rpi = EW_THREADS / ldv;
  const int g = blockIdx.y;
  const int cg = threadIdx.x % ldv, rs = threadIdx.x / ldv;
  double s[VEC], q[VEC];  // fp64 accumulation (B200 runs fp64 adds at half the fp32 rate; the kernel is HBM bound)
#pragma unroll
  for (int j = 0; j < VEC; ++j) s[j] = q[j] = 0.0;
  if (rs < rpi) {
    for (long long r = (long long)blockIdx.x * rpi + rs; r < R; r += (long long)gridDim.x * rpi) {
      const float* p = x + ((long long)g * R + r) * ld + cg * VEC;
      float v[VEC];
      if constexpr (VEC == 4) {
        const float4 t = ldg4(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        v[0] = __ldg(p);
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) { s[j] += (double)v[j]; q[j] += (double)v[j] * (double)v[j]; }
    }
  }
  const int W = ldv * VEC;
  if (rs < rpi) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      red[(rs * 2 + 0) * W + cg * VEC + j] = s[j];
      red[(rs * 2 + 1) * W + cg * VEC + j] = q[j];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * W; idx += EW_THREADS) {
    const int which = idx / W, col = idx % W;
    if (col < C) {
      double t = 0.0;
      for (int r = 0; r < rpi; ++r) t += red[(r * 2 + which) * W + col];
      atomicAdd(stats + ((long long)g * 2 + which) * C + col, t);
    }
  }
}

extern "C" int sb_col_stats(const float* x, int64_t ld, int64_t R, int32_t G, int32_t C, double* stats,
                            void* stream) {
  SB_CHECK_ARG(R >= 0 && G >= 1 && C >= 1 && ld >= C && ld <= 1024, "sb_col_stats: bad sizes");
  if (R == 0) return SB_OK;
  const bool vec = (ld % 4 == 0) && ((uintptr_t)x % 16 == 0);
  const int VEC = vec ? 4 : 1;
  const int ldv = (int)(ld / VEC);
  SB_CHECK_ARG(ldv <= EW_THREADS, "sb_col_stats: row too wide");
  const int rpi = EW_THREADS / ldv;
  long long blocks = sb_ceil_div(R, (long long)rpi * 8);
  const long long cap = (long long)sb_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)G);
  const size_t smem = (size_t)rpi * 2 * ldv * VEC * sizeof(double);
  if (vec) col_stats_kernel<4><<<grid, EW_THREADS, smem, (cudaStream_t)stream>>>(x, ld, R, C, stats);
  else col_stats_kernel<1><<<grid, EW_THREADS, smem, (cudaStream_t)stream>>>(x, ld, R, C, stats);
  SB_CHECK_LAUNCH("sb_col_stats");
  return SB_OK;
}

// ------------------------------------------------------------------------------------- forward: affine(+relu)(+res)
// out[g, r, c] = act(a[g,c] * y[g,r,c] + c[g,c]) + res[g,r,c]      (pad columns C..ld-1 -> 0)
__global__ void __launch_bounds__(EW_THREADS) affine_act_res_kernel(const float* __restrict__ y,
                                                                    const float* __restrict__ pa,
                                                                    const float* __restrict__ pc,
                                                                    const float* __restrict__ res,
                                                                    float* __restrict__ out, long long ld,
                                                                    long long R, int G, int C, int relu) {
  const long long ld4 = ld >> 2;
  const long long total = (long long)G * R * ld4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / ld4;
    const int c4 = (int)(t - row * ld4);
    const int g = (int)(row / R);
    const float4 v = ldg4(y + t * 4);
    float in[4] = {v.x, v.y, v.z, v.w}, o[4];
    float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) rr = ldg4(res + t * 4);
    const float rv[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = c4 * 4 + j;
      if (col < C) {
        float u = pa ? fmaf(__ldg(pa + (long long)g * C + col), in[j], __ldg(pc + (long long)g * C + col)) : in[j];
        if (relu) u = fmaxf(u, 0.f);
        o[j] = u + rv[j];
      } else {
        o[j] = 0.f;
      }
    }
    *reinterpret_cast<float4*>(out + t * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

extern "C" int sb_affine_act_res(const float* y, const float* pa, const float* pc, const float* res, float* out,
                                 int64_t ld, int64_t R, int32_t G, int32_t C, int32_t relu, void* stream) {
  SB_CHECK_ARG(ld % 4 == 0 && ld >= C && G >= 1, "sb_affine_act_res: ld must be a multiple of 4 and >= C");
  SB_CHECK_ARG((pa == nullptr) == (pc == nullptr), "sb_affine_act_res: pa/pc must come together");
  if (R == 0) return SB_OK;
  const long long total = (long long)G * R * (ld / 4);
  long long blocks = sb_ceil_div(total, EW_THREADS * 4);
  const long long cap = (long long)sb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  affine_act_res_kernel<<<(unsigned)blocks, EW_THREADS, 0, (cudaStream_t)stream>>>(y, pa, pc, res, out, ld, R, G, C,
                                                                                   relu);
  SB_CHECK_LAUNCH("sb_affine_act_res");
  return SB_OK;
}

// --------------------------------------------------------------------------- backward: relu mask + BN reductions
// dz = gout * [a*y + c > 0]  (relu) or gout;  s1[g,c] += sum_r dz ;  s2[g,c] += sum_r dz * (y - mean) * rstd.
// dz may alias gout (in place).
__global__ void __launch_bounds__(EW_THREADS, 3) bn_bwd_reduce_kernel(const float* gout, const float* __restrict__ y,
                                                                   const float* __restrict__ pa,
                                                                   const float* __restrict__ pc,
                                                                   const double* __restrict__ mean_rstd, float* dz,
                                                                   long long ld, long long R, int G, int C, int relu,
                                                                   double* __restrict__ stats) {
  extern __shared__ double red[];
  const int ldv = (int)(ld >> 2);
  const int rpi = EW_THREADS / ldv;
  const int g = blockIdx.y;
  const int cg = threadIdx.x % ldv, rs = threadIdx.x / ldv;
  // fp64 accumulation: the BatchNorm backward projects out mean and x_hat components, and gradients upstream are
  // the small residual of that cancellation (amplified by var/eps), so the sums must be far more accurate than fp32.
  double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
  if (rs < rpi) {
    float a4[4], c4[4];
    double m4[4], r4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = cg * 4 + j;
      const bool ok = col < C;
      a4[j] = ok ? __ldg(pa + (long long)g * C + col) : 0.f;
      c4[j] = ok ? __ldg(pc + (long long)g * C + col) : 0.f;
      m4[j] = ok ? mean_rstd[(long long)g * C + col] : 0.0;
      r4[j] = ok ? mean_rstd[((long long)G + g) * C + col] : 0.0;
    }
    const long long rstep = (long long)gridDim.x * rpi;
    // two rows per thread and iteration in flight (pure stream: bytes in flight are what buys HBM bandwidth)
    for (long long r0 = (long long)blockIdx.x * rpi + rs; r0 < R; r0 += 2 * rstep) {
      const long long rr[2] = {r0, r0 + rstep};
      float4 gv2[2], yv2[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (rr[i] < R) {
          const long long off = ((long long)g * R + rr[i]) * ld + cg * 4;
          gv2[i] = *reinterpret_cast<const float4*>(gout + off);
          yv2[i] = ldg4(y + off);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (rr[i] >= R) break;
        const long long off = ((long long)g * R + rr[i]) * ld + cg * 4;
        const float gi[4] = {gv2[i].x, gv2[i].y, gv2[i].z, gv2[i].w}, yi[4] = {yv2[i].x, yv2[i].y, yv2[i].z, yv2[i].w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d = gi[j];
          if (relu && !(fmaf(a4[j], yi[j], c4[j]) > 0.f)) d = 0.f;
          if (cg * 4 + j >= C) d = 0.f;
          o[j] = d;
          s1[j] += (double)d;
          s2[j] += (double)d * (((double)yi[j] - m4[j]) * r4[j]);
        }
        if (dz) *reinterpret_cast<float4*>(dz + off) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  const int W = ldv * 4;
  if (rs < rpi) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[(rs * 2 + 0) * W + cg * 4 + j] = s1[j];
      red[(rs * 2 + 1) * W + cg * 4 + j] = s2[j];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * W; idx += EW_THREADS) {
    const int which = idx / W, col = idx % W;
    if (col < C) {
      double t = 0.0;
      for (int r = 0; r < rpi; ++r) t += red[(r * 2 + which) * W + col];
      atomicAdd(stats + ((long long)g * 2 + which) * C + col, t);
    }
  }
}

extern "C" int sb_bn_bwd_reduce(const float* gout, const float* y, const float* pa, const float* pc,
                                const double* mean_rstd, float* dz, int64_t ld, int64_t R, int32_t G, int32_t C,
                                int32_t relu, double* stats, void* stream) {
  SB_CHECK_ARG(ld % 4 == 0 && ld >= C && ld / 4 <= EW_THREADS && G >= 1, "sb_bn_bwd_reduce: bad leading dim");
  SB_CHECK_ARG(gout && y && pa && pc && mean_rstd && stats, "sb_bn_bwd_reduce: null argument");
  if (R == 0) return SB_OK;
  const int ldv = (int)(ld / 4);
  const int rpi = EW_THREADS / ldv;
  long long blocks = sb_ceil_div(R, (long long)rpi * 8);
  const long long cap = (long long)sb_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)G);
  const size_t smem = (size_t)rpi * 2 * ldv * 4 * sizeof(double);
  bn_bwd_reduce_kernel<<<grid, EW_THREADS, smem, (cudaStream_t)stream>>>(gout, y, pa, pc, mean_rstd, dz, ld, R, G, C,
                                                                        relu, stats);
  SB_CHECK_LAUNCH("sb_bn_bwd_reduce");
  return SB_OK;
}

// dgamma (+)= sum_g s2[g], dbeta (+)= sum_g s1[g]; fp64 coefficients of the centred form
//     dY = al * dZ + be * (Y - mean) + ga
//   training:  al = a, be = -a * rstd^2 * m2, ga = -a * m1        (a = gamma * rstd, m1 = s1/M, m2 = s2/M)
//   eval:      al = a, be = 0,                ga = 0
// Centred + fp64 on purpose: the upstream weight gradient is the eps/(var+eps)-sized residual of
// sum dY * Y, so a coefficient rounded to fp32 (relative 6e-8) shows up amplified by var/eps (and by mean^2/eps in the
// uncentred form) in dW of the preceding Linear.
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ stats, long long M, int G, int C,
                                       const float* __restrict__ gamma, const double* __restrict__ mean_rstd,
                                       int training, int accumulate, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, double* __restrict__ coef /*[3,G,C]*/) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  double dg = 0.0, db = 0.0;
  const double gm = gamma ? (double)gamma[ch] : 1.0;
  for (int g = 0; g < G; ++g) {
    const double s1 = stats[((long long)g * 2 + 0) * C + ch], s2 = stats[((long long)g * 2 + 1) * C + ch];
    dg += s2;
    db += s1;
    const double rs = mean_rstd[((long long)G + g) * C + ch];
    const double a = gm * rs;
    double bev = 0.0, gav = 0.0;
    if (training) {
      const double m1 = s1 / (double)M, m2 = s2 / (double)M;
      bev = -a * rs * m2;
      gav = -a * m1;
    }
    coef[((long long)0 * G + g) * C + ch] = a;
    coef[((long long)1 * G + g) * C + ch] = bev;
    coef[((long long)2 * G + g) * C + ch] = gav;
  }
  if (dgamma) dgamma[ch] = accumulate ? dgamma[ch] + (float)dg : (float)dg;
  if (dbeta) dbeta[ch] = accumulate ? dbeta[ch] + (float)db : (float)db;
}

extern "C" int sb_bn_bwd_finalize(const double* stats, int64_t M, int32_t G, int32_t C, const float* gamma,
                                  const double* mean_rstd, int32_t training, int32_t accumulate, float* dgamma,
                                  float* dbeta, double* coef, void* stream) {
  SB_CHECK_ARG(stats && mean_rstd && coef && M >= 1, "sb_bn_bwd_finalize: null argument");
  bn_bwd_finalize_kernel<<<(unsigned)sb_ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(
      stats, M, G, C, gamma, mean_rstd, training, accumulate, dgamma, dbeta, coef);
  SB_CHECK_LAUNCH("sb_bn_bwd_finalize");
  return SB_OK;
}

// out = al[g,c]*t1 + be[g,c]*(t2 - mean[g,c]) + ga[g,c]   (out may alias t1).
// The fp64 coefficients are carried as unevaluated float pairs (hi + lo) and applied with two FMAs each, so the
// coefficient itself contributes no systematic fp32 rounding (see bn_bwd_finalize_kernel) while the kernel stays an
// HBM-bound fp32 stream; t2 - mean is formed against the float pair of the fp64 mean.
struct F2 { float hi, lo; };
__device__ __forceinline__ F2 split_d(double v) {
  F2 r;
  r.hi = (float)v;
  r.lo = (float)(v - (double)r.hi);
  return r;
}
__global__ void __launch_bounds__(EW_THREADS, 4) affine2_kernel(const float* t1, const float* __restrict__ t2,
                                                             const double* __restrict__ coef,
                                                             const double* __restrict__ mean_rstd,
                                                             const float* __restrict__ pa,
                                                             const float* __restrict__ pc, float* out, long long ld,
                                                             long long R, int G, int C,
                                                             const double* __restrict__ fin_stats = nullptr,
                                                             const float* __restrict__ gamma = nullptr,
                                                             long long M = 0, int training = 0,
                                                             float* __restrict__ dgamma = nullptr,
                                                             float* __restrict__ dbeta = nullptr) {
  // tab[k][g*ld + col], k = (al.hi, al.lo, be.hi, be.lo, ga.hi, ga.lo, mu.hi, mu.lo, pa, pc): fp64 -> float pairs once
  // per block (not per element); structure-of-arrays so a warp's float4 reads are conflict-free.
  extern __shared__ __align__(16) float tab[];
  const long long GC = (long long)G * C;
  const int GL = G * (int)ld;
  const bool mask = pa != nullptr;   // t1 is the upstream gradient: apply the ReLU mask [pa*t2 + pc > 0] here
  for (int p = threadIdx.x; p < GL; p += blockDim.x) {
    const int g = p / (int)ld, col = p - g * (int)ld;
    F2 al = {0.f, 0.f}, be = al, ga = al, mu = al;
    float a_ = 0.f, c_ = 0.f;
    if (col < C) {
      const long long q = (long long)g * C + col;
      if (fin_stats) {   // sb_bn_apply_bwd: the coefficients straight from the two column sums, the expressions of
                         // bn_bwd_finalize_kernel (identical values), instead of a 4 us launch of their own
        const double s1 = fin_stats[((long long)g * 2 + 0) * C + col], s2 = fin_stats[((long long)g * 2 + 1) * C + col];
        const double gm = gamma ? (double)gamma[col] : 1.0;
        const double rs = mean_rstd[((long long)G + g) * C + col];
        const double av = gm * rs;
        double bev = 0.0, gav = 0.0;
        if (training) {
          const double m1 = s1 / (double)M, m2 = s2 / (double)M;
          bev = -av * rs * m2;
          gav = -av * m1;
        }
        al = split_d(av); be = split_d(bev); ga = split_d(gav);
      } else {
        al = split_d(coef[q]); be = split_d(coef[GC + q]); ga = split_d(coef[2 * GC + q]);
      }
      mu = split_d(mean_rstd[q]);
      if (mask) { a_ = __ldg(pa + q); c_ = __ldg(pc + q); }
    }
    tab[0 * GL + p] = al.hi; tab[1 * GL + p] = al.lo; tab[2 * GL + p] = be.hi; tab[3 * GL + p] = be.lo;
    tab[4 * GL + p] = ga.hi; tab[5 * GL + p] = ga.lo; tab[6 * GL + p] = mu.hi; tab[7 * GL + p] = mu.lo;
    tab[8 * GL + p] = a_; tab[9 * GL + p] = c_;
  }
  if (fin_stats && blockIdx.x == 0) {   // dgamma, dbeta: sums over the groups in order, as bn_bwd_finalize_kernel
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      double dg = 0.0, db = 0.0;
      for (int g = 0; g < G; ++g) {
        dg += fin_stats[((long long)g * 2 + 1) * C + ch];
        db += fin_stats[((long long)g * 2 + 0) * C + ch];
      }
      if (dgamma) dgamma[ch] = (float)dg;
      if (dbeta) dbeta[ch] = (float)db;
    }
  }
  __syncthreads();
  const long long ld4 = ld >> 2;
  const long long total = (long long)G * R * ld4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // two independent float4 items per thread and iteration: 64 B of loads in flight per thread (the kernel is a pure
  // stream; with one item it sat at ~70 % of the HBM peak for lack of bytes in flight at 3 CTAs / SM)
  for (long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; t0 < total; t0 += 2 * stride) {
    const long long tt[2] = {t0, t0 + stride};
    float4 uu[2], vv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (tt[i] < total) {
        uu[i] = *reinterpret_cast<const float4*>(t1 + tt[i] * 4);
        vv[i] = ldg4(t2 + tt[i] * 4);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (tt[i] >= total) break;
      const long long t = tt[i];
      const long long row = t / ld4;
      const int c4 = (int)(t - row * ld4);
      const int g = (row >= R) ? (int)(row / R) : 0;
      float4 u = uu[i];
      const float4 v = vv[i];
      const int p = g * (int)ld + c4 * 4;
      float4 k[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) k[q] = *reinterpret_cast<const float4*>(tab + q * GL + p);
      if (mask) {   // same predicate as bn_bwd_reduce_kernel: dz = gout * [pa*y + pc > 0]
        const float4 a4 = *reinterpret_cast<const float4*>(tab + 8 * GL + p);
        const float4 c4v = *reinterpret_cast<const float4*>(tab + 9 * GL + p);
        if (!(fmaf(a4.x, v.x, c4v.x) > 0.f)) u.x = 0.f;
        if (!(fmaf(a4.y, v.y, c4v.y) > 0.f)) u.y = 0.f;
        if (!(fmaf(a4.z, v.z, c4v.z) > 0.f)) u.z = 0.f;
        if (!(fmaf(a4.w, v.w, c4v.w) > 0.f)) u.w = 0.f;
      }
      float4 o;
#define SB_AFF2(c)                                                                 \
      {                                                                            \
        const float d = (v.c - k[6].c) - k[7].c;      /* t2 - mean */              \
        float acc = fmaf(k[3].c, d, k[5].c);          /* be.lo * d + ga.lo */      \
        acc = fmaf(k[1].c, u.c, acc);                 /* + al.lo * t1 */           \
        acc = fmaf(k[2].c, d, acc + k[4].c);          /* + be.hi * d + ga.hi */    \
        o.c = fmaf(k[0].c, u.c, acc);                 /* + al.hi * t1 */           \
      }
      SB_AFF2(x) SB_AFF2(y) SB_AFF2(z) SB_AFF2(w)
#undef SB_AFF2
      // padding columns: every table entry is 0 there, so o = 0 as required
      *reinterpret_cast<float4*>(out + t * 4) = o;
    }
  }
}

extern "C" int sb_affine2(const float* t1, const float* t2, const double* coef, const double* mean_rstd,
                          const float* pa, const float* pc, float* out, int64_t ld, int64_t R, int32_t G, int32_t C,
                          void* stream) {
  SB_CHECK_ARG(ld % 4 == 0 && ld >= C && G >= 1, "sb_affine2: ld must be a multiple of 4 and >= C");
  SB_CHECK_ARG((pa == nullptr) == (pc == nullptr), "sb_affine2: pa/pc must come together");
  if (R == 0) return SB_OK;
  const long long total = (long long)G * R * (ld / 4);
  long long blocks = sb_ceil_div(total, EW_THREADS * 4);
  const long long cap = (long long)sb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)G * ld * 10 * sizeof(float);
  SB_CHECK_ARG(smem <= 48 * 1024, "sb_affine2: G*ld too large");
  affine2_kernel<<<(unsigned)blocks, EW_THREADS, smem, (cudaStream_t)stream>>>(t1, t2, coef, mean_rstd, pa, pc, out, ld,
                                                                              R, G, C);
  SB_CHECK_LAUNCH("sb_affine2");
  return SB_OK;
}

// sb_bn_bwd_finalize + sb_affine2 in ONE launch: dz = al*dZ' + be*(y - mean) + ga with dZ' = gout * [pa*y + pc > 0]
// (pa/pc given) and the coefficients taken straight from stats[G,2,C] = (sum dZ', sum dZ' y_hat); writes dgamma, dbeta.
extern "C" int sb_bn_apply_bwd(const float* gout, const float* y, const double* stats, const double* mean_rstd,
                               const float* pa, const float* pc, const float* gamma, int64_t M, int32_t training,
                               float* dz, float* dgamma, float* dbeta, int64_t ld, int64_t R, int32_t G, int32_t C,
                               void* stream) {
  SB_CHECK_ARG(gout && y && stats && mean_rstd && dz && ld % 4 == 0 && ld >= C && G >= 1 && M >= 1,
               "sb_bn_apply_bwd: bad arguments");
  SB_CHECK_ARG((pa == nullptr) == (pc == nullptr), "sb_bn_apply_bwd: pa/pc must come together");
  const long long total = (long long)G * R * (ld / 4);
  long long blocks = sb_ceil_div(total, EW_THREADS * 4);
  const long long cap = (long long)sb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;   // R == 0 still writes dgamma / dbeta
  const size_t smem = (size_t)G * ld * 10 * sizeof(float);
  SB_CHECK_ARG(smem <= 48 * 1024, "sb_bn_apply_bwd: G*ld too large");
  affine2_kernel<<<(unsigned)blocks, EW_THREADS, smem, (cudaStream_t)stream>>>(gout, y, nullptr, mean_rstd, pa, pc, dz, ld,
                                                                              R, G, C, stats, gamma, M, training, dgamma,
                                                                              dbeta);
  SB_CHECK_LAUNCH("sb_bn_apply_bwd");
  return SB_OK;
}

// ------------------------------------------------------------- BatchNorm finalize + apply in ONE launch (forward)
// sb_bn_finalize followed by sb_affine_act_res costs two launches, the first of them a 5 us kernel that only turns
// 2*G*C sums into coefficients.  Here every CTA derives the coefficients of all (group, channel) pairs itself (a few
// hundred fp64 operations, the same expressions as bn_finalize_kernel -> identical a, c) into shared memory, and CTA 0
// also publishes a, c, mean_rstd for the backward and updates the running buffers (+v then -v order, as before).
__global__ void __launch_bounds__(EW_THREADS) bn_apply_fwd_kernel(
    const float* __restrict__ y, const double* __restrict__ stats, long long M, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* running_mean, float* running_var, float momentum, float eps, int training,
    const float* __restrict__ res, float* __restrict__ out, long long ld, long long R, int G, int C, int relu,
    float* __restrict__ a_out, float* __restrict__ c_out, double* __restrict__ mean_rstd) {
  extern __shared__ __align__(16) float tab[];   // [2][G*ld]: a, c (0 in padding columns)
  const int GL = G * (int)ld;
  for (int p = threadIdx.x; p < GL; p += blockDim.x) {
    const int g = p / (int)ld, ch = p - g * (int)ld;
    float av_f = 0.f, cv_f = 0.f;
    if (ch < C) {
      const double gm = gamma ? (double)gamma[ch] : 1.0, bt = beta ? (double)beta[ch] : 0.0;
      double mean, var;
      if (training) {
        const double sv = stats[((long long)g * 2 + 0) * C + ch], q = stats[((long long)g * 2 + 1) * C + ch];
        mean = sv / (double)M;
        var = q / (double)M - mean * mean;
        if (var < 0.0) var = 0.0;
      } else {
        mean = (double)running_mean[ch];
        var = (double)running_var[ch];
      }
      const double rstd = 1.0 / sqrt(var + (double)eps);
      const double av = gm * rstd;
      av_f = (float)av;
      cv_f = (float)(bt - mean * av);
      if (blockIdx.x == 0) {
        a_out[(long long)g * C + ch] = av_f;
        c_out[(long long)g * C + ch] = cv_f;
        if (mean_rstd) {
          mean_rstd[(long long)g * C + ch] = mean;
          mean_rstd[((long long)G + g) * C + ch] = rstd;
        }
      }
    }
    tab[p] = av_f;
    tab[GL + p] = cv_f;
  }
  if (blockIdx.x == 0 && training && running_mean) {   // one thread per channel, groups in order (+v, then -v)
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      for (int g = 0; g < G; ++g) {
        const double sv = stats[((long long)g * 2 + 0) * C + ch], q = stats[((long long)g * 2 + 1) * C + ch];
        const double mean = sv / (double)M;
        double var = q / (double)M - mean * mean;
        if (var < 0.0) var = 0.0;
        const double unb = (M > 1) ? var * ((double)M / (double)(M - 1)) : var;
        running_mean[ch] = (float)((1.0 - (double)momentum) * (double)running_mean[ch] + (double)momentum * mean);
        running_var[ch] = (float)((1.0 - (double)momentum) * (double)running_var[ch] + (double)momentum * unb);
      }
    }
  }
  __syncthreads();
  const long long ld4 = ld >> 2;
  const long long total = (long long)G * R * ld4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / ld4;
    const int c4 = (int)(t - row * ld4);
    const int g = (int)(row / R);
    const float4 v = ldg4(y + t * 4);
    float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) rr = ldg4(res + t * 4);
    const float4 a4 = *reinterpret_cast<const float4*>(tab + g * (int)ld + c4 * 4);
    const float4 k4 = *reinterpret_cast<const float4*>(tab + GL + g * (int)ld + c4 * 4);
    const float in[4] = {v.x, v.y, v.z, v.w}, rv[4] = {rr.x, rr.y, rr.z, rr.w};
    const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, cc[4] = {k4.x, k4.y, k4.z, k4.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // same arithmetic as affine_act_res_kernel
      if (c4 * 4 + j < C) {
        float u = fmaf(aa[j], in[j], cc[j]);
        if (relu) u = fmaxf(u, 0.f);
        o[j] = u + rv[j];
      } else {
        o[j] = 0.f;
      }
    }
    *reinterpret_cast<float4*>(out + t * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

extern "C" int sb_bn_apply_fwd(const float* y, const double* stats, int64_t M, int32_t G, int32_t C, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                               int32_t training, int32_t relu, const float* res, float* out, int64_t ld, int64_t R,
                               float* a, float* c, double* mean_rstd, void* stream) {
  SB_CHECK_ARG(y && out && a && c && G >= 1 && C >= 1 && ld % 4 == 0 && ld >= C, "sb_bn_apply_fwd: bad arguments");
  SB_CHECK_ARG(training ? (stats != nullptr && M >= 1) : (running_mean && running_var),
               "sb_bn_apply_fwd: training needs stats and M>=1, eval needs running statistics");
  const size_t smem = (size_t)2 * G * ld * sizeof(float);
  if (smem > 40 * 1024) {   // very wide rows: the two-launch form
    const int rc = sb_bn_finalize(stats, M, G, C, gamma, beta, running_mean, running_var, momentum, eps, training, a, c,
                                  mean_rstd, stream);
    if (rc) return rc;
    return sb_affine_act_res(y, a, c, res, out, ld, R, G, C, relu, stream);
  }
  const long long total = (long long)G * R * (ld / 4);
  long long blocks = sb_ceil_div(total, EW_THREADS * 4);
  const long long cap = (long long)sb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;   // R == 0 still publishes the coefficients and moves the running buffers
  bn_apply_fwd_kernel<<<(unsigned)blocks, EW_THREADS, smem, (cudaStream_t)stream>>>(
      y, stats, M, gamma, beta, running_mean, running_var, momentum, eps, training, res, out, ld, R, G, C, relu, a, c,
      mean_rstd);
  SB_CHECK_LAUNCH("sb_bn_apply_fwd");
  return SB_OK;
}

// ------------------------------------------------------------------------------ slot / sign reduction for rho
// out[node, c] = sum_s sum_{j < k_b} x[s, row(b, j, i), c]      ("sum over the k eigenvector slots and both signs",
// sign_net.py:113 + :70 / deepsigns.py:72-81).  One warp per node, fixed summation order (s outer, j inner).
__global__ void __launch_bounds__(256) slot_sum_fwd_kernel(const float* __restrict__ x, long long ld, long long R,
                                                           int S, const int64_t* __restrict__ batch,
                                                           const int32_t* __restrict__ gp,
                                                           const int64_t* __restrict__ row_ptr, long long N, int k,
                                                           int masked, int k_limit_by_n, float* __restrict__ out,
                                                           long long ldo, int C) {
  const long long node = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= N) return;
  const int b = (int)batch[node];
  const int n = gp[b + 1] - gp[b];
  int kb = masked ? (n < k ? n : k) : k;
  if (k_limit_by_n && kb > n) kb = n;  // DGL masked variant: slots >= n_b are zeroed before the sum
  const int li = (int)(node - gp[b]);
  const long long r0 = row_ptr[b] + li;
  for (int c = lane; c < (int)ldo; c += 32) {
    float acc = 0.f;
    if (c < C) {
      for (int s = 0; s < S; ++s)
        for (int j = 0; j < kb; ++j) acc += __ldg(x + ((long long)s * R + r0 + (long long)j * n) * ld + c);
    }
    out[node * ldo + c] = acc;
  }
}
extern "C" int sb_slot_sum_fwd(const float* x, int64_t ld, int64_t R, int32_t S, const int64_t* batch,
                               const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N, int32_t k,
                               int32_t masked, int32_t limit_by_n, float* out, int64_t ldo, int32_t C,
                               void* stream) {
  if (N == 0) return SB_OK;
  slot_sum_fwd_kernel<<<(unsigned)sb_ceil_div(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      x, ld, R, S, batch, graph_ptr, row_ptr, N, k, masked, limit_by_n, out, ldo, C);
  SB_CHECK_LAUNCH("sb_slot_sum_fwd");
  return SB_OK;
}

// backward: gx[s, row(b,j,i), c] = gout[node, c] for contributing slots (0 for slots zeroed by limit_by_n)
__global__ void __launch_bounds__(256) slot_sum_bwd_kernel(const float* __restrict__ gout, long long ldo,
                                                           float* __restrict__ gx, long long ld, long long R, int S,
                                                           const int64_t* __restrict__ batch,
                                                           const int32_t* __restrict__ gp,
                                                           const int64_t* __restrict__ row_ptr, long long N, int k,
                                                           int masked, int k_limit_by_n, int C) {
  const long long node = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= N) return;
  const int b = (int)batch[node];
  const int n = gp[b + 1] - gp[b];
  const int kb = masked ? (n < k ? n : k) : k;
  int klive = kb;
  if (k_limit_by_n && klive > n) klive = n;
  const int li = (int)(node - gp[b]);
  const long long r0 = row_ptr[b] + li;
  for (int c = lane; c < (int)ld; c += 32) {
    const float gv = (c < C) ? __ldg(gout + node * ldo + c) : 0.f;
    for (int s = 0; s < S; ++s)
      for (int j = 0; j < kb; ++j) gx[((long long)s * R + r0 + (long long)j * n) * ld + c] = (j < klive) ? gv : 0.f;
  }
}
extern "C" int sb_slot_sum_bwd(const float* gout, int64_t ldo, float* gx, int64_t ld, int64_t R, int32_t S,
                               const int64_t* batch, const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N,
                               int32_t k, int32_t masked, int32_t limit_by_n, int32_t C, void* stream) {
  if (N == 0) return SB_OK;
  slot_sum_bwd_kernel<<<(unsigned)sb_ceil_div(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      gout, ldo, gx, ld, R, S, batch, graph_ptr, row_ptr, N, k, masked, limit_by_n, C);
  SB_CHECK_LAUNCH("sb_slot_sum_bwd");
  return SB_OK;
}

// out = g * [y > 0]   (backward of a ReLU fused into a Linear epilogue; out may alias g)
__global__ void relu_bwd_kernel(const float* g, const float* __restrict__ y, float* out, long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    out[t] = (__ldg(y + t) > 0.f) ? g[t] : 0.f;
}
extern "C" int sb_relu_bwd(const float* g, const float* y, float* out, int64_t n, void* stream) {
  if (n == 0) return SB_OK;
  long long blocks = sb_ceil_div(n, 256 * 4);
  const long long cap = (long long)sb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  relu_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, y, out, n);
  SB_CHECK_LAUNCH("sb_relu_bwd");
  return SB_OK;
}

// =====================================================================================================================
// BatchNorm + activation as ONE call (and, for small tensors, ONE kernel).
// The predictor / rho / encoder side of the path normalises [N, d] tensors with a few thousand rows: three launches
// (column sums, finalize, apply) of ~4 us each per BatchNorm forward and four per backward are pure launch latency there
// (cfg 2, the reference's Alchemy configuration: ~1100 launches for a step of ~5 ms of GPU work).  For M <= 8192 rows and
// one group a CTA owns four channels end to end: fp64 column sums over all rows, the affine coefficients (+ running
// buffers), then the apply pass over the same rows (L2-resident).  Larger tensors take the streaming kernels above.
#define BN_SMALL_MAX_ROWS 8192

__device__ __forceinline__ void bn_small_reduce8(double (&v)[8], double* red /*[8 warps][8]*/) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = warp_sum_d(v[j]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp * 8 + j] = v[j];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double t = 0.0;
    for (int w = 0; w < EW_THREADS / 32; ++w) t += red[w * 8 + threadIdx.x];
    red[64 + threadIdx.x] = t;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = red[64 + j];
}

__global__ void __launch_bounds__(EW_THREADS) bn_act_small_fwd_kernel(
    const float* __restrict__ x, long long ld, long long M, int C, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
    float eps, int training, int relu, const float* __restrict__ res, float* __restrict__ out, float* __restrict__ a,
    float* __restrict__ c, double* __restrict__ mean_rstd) {
  __shared__ double red[72];
  __shared__ float sa[4], sc[4];
  const int col0 = blockIdx.x * 4;
  double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (training) {
    for (long long r = threadIdx.x; r < M; r += EW_THREADS) {
      const float4 t = ldg4(x + r * ld + col0);
      v[0] += (double)t.x; v[1] += (double)t.y; v[2] += (double)t.z; v[3] += (double)t.w;
      v[4] += (double)t.x * (double)t.x; v[5] += (double)t.y * (double)t.y;
      v[6] += (double)t.z * (double)t.z; v[7] += (double)t.w * (double)t.w;
    }
    bn_small_reduce8(v, red);
  }
  if (threadIdx.x < 4) {
    const int ch = col0 + threadIdx.x;
    float av = 0.f, cv = 0.f;
    if (ch < C) {   // same arithmetic as bn_finalize_kernel
      const double gm = gamma ? (double)gamma[ch] : 1.0, bt = beta ? (double)beta[ch] : 0.0;
      double mean, var;
      if (training) {
        mean = red[64 + threadIdx.x] / (double)M;
        var = red[68 + threadIdx.x] / (double)M - mean * mean;
        if (var < 0.0) var = 0.0;
        if (running_mean) {
          const double unb = (M > 1) ? var * ((double)M / (double)(M - 1)) : var;
          running_mean[ch] = (float)((1.0 - (double)momentum) * (double)running_mean[ch] + (double)momentum * mean);
          running_var[ch] = (float)((1.0 - (double)momentum) * (double)running_var[ch] + (double)momentum * unb);
        }
      } else {
        mean = (double)running_mean[ch];
        var = (double)running_var[ch];
      }
      const double rstd = 1.0 / sqrt(var + (double)eps);
      const double avd = gm * rstd;
      av = (float)avd;
      cv = (float)(bt - mean * avd);
      a[ch] = av;
      c[ch] = cv;
      if (mean_rstd) { mean_rstd[ch] = mean; mean_rstd[C + ch] = rstd; }
    }
    sa[threadIdx.x] = av;
    sc[threadIdx.x] = cv;
  }
  __syncthreads();
  const float a4[4] = {sa[0], sa[1], sa[2], sa[3]}, c4[4] = {sc[0], sc[1], sc[2], sc[3]};
  for (long long r = threadIdx.x; r < M; r += EW_THREADS) {
    const float4 t = ldg4(x + r * ld + col0);
    float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) rr = ldg4(res + r * ld + col0);
    const float in[4] = {t.x, t.y, t.z, t.w}, rv[4] = {rr.x, rr.y, rr.z, rr.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // same arithmetic as affine_act_res_kernel
      if (col0 + j < C) {
        float u = fmaf(a4[j], in[j], c4[j]);
        if (relu) u = fmaxf(u, 0.f);
        o[j] = u + rv[j];
      } else {
        o[j] = 0.f;
      }
    }
    *reinterpret_cast<float4*>(out + r * ld + col0) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(EW_THREADS) bn_act_small_bwd_kernel(
    const float* gout, const float* __restrict__ x, const float* __restrict__ pa, const float* __restrict__ pc,
    const double* __restrict__ mean_rstd, const float* __restrict__ gamma, long long ld, long long M, int C, int relu,
    int training, float* dz, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double red[72];
  __shared__ float tab[8][4];   // al.hi, al.lo, be.hi, be.lo, ga.hi, ga.lo, mu.hi, mu.lo per channel
  const int col0 = blockIdx.x * 4;
  float a4[4], c4[4];
  double m4[4], r4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool ok = col0 + j < C;
    a4[j] = ok ? __ldg(pa + col0 + j) : 0.f;
    c4[j] = ok ? __ldg(pc + col0 + j) : 0.f;
    m4[j] = ok ? mean_rstd[col0 + j] : 0.0;
    r4[j] = ok ? mean_rstd[C + col0 + j] : 0.0;
  }
  double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (long long r = threadIdx.x; r < M; r += EW_THREADS) {   // same arithmetic as bn_bwd_reduce_kernel
    const float4 g4 = *reinterpret_cast<const float4*>(gout + r * ld + col0);
    const float4 y4 = ldg4(x + r * ld + col0);
    const float gi[4] = {g4.x, g4.y, g4.z, g4.w}, yi[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float d = gi[j];
      if (relu && !(fmaf(a4[j], yi[j], c4[j]) > 0.f)) d = 0.f;
      if (col0 + j >= C) d = 0.f;
      v[j] += (double)d;
      v[4 + j] += (double)d * (((double)yi[j] - m4[j]) * r4[j]);
    }
  }
  bn_small_reduce8(v, red);
  if (threadIdx.x < 4) {   // same arithmetic as bn_bwd_finalize_kernel (G = 1)
    const int ch = col0 + threadIdx.x;
    F2 al = {0.f, 0.f}, be = al, ga = al, mu = al;
    if (ch < C) {
      const double s1 = red[64 + threadIdx.x], s2 = red[68 + threadIdx.x];
      const double gm = gamma ? (double)gamma[ch] : 1.0;
      const double rs = mean_rstd[C + ch];
      const double av = gm * rs;
      double bev = 0.0, gav = 0.0;
      if (training) {
        const double m1 = s1 / (double)M, m2 = s2 / (double)M;
        bev = -av * rs * m2;
        gav = -av * m1;
      }
      if (dgamma) dgamma[ch] = (float)s2;
      if (dbeta) dbeta[ch] = (float)s1;
      al = split_d(av); be = split_d(bev); ga = split_d(gav); mu = split_d(mean_rstd[ch]);
    }
    tab[0][threadIdx.x] = al.hi; tab[1][threadIdx.x] = al.lo; tab[2][threadIdx.x] = be.hi; tab[3][threadIdx.x] = be.lo;
    tab[4][threadIdx.x] = ga.hi; tab[5][threadIdx.x] = ga.lo; tab[6][threadIdx.x] = mu.hi; tab[7][threadIdx.x] = mu.lo;
  }
  __syncthreads();
  float k[8][4];
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) k[q][j] = tab[q][j];
  for (long long r = threadIdx.x; r < M; r += EW_THREADS) {   // same arithmetic as affine2_kernel
    const float4 g4 = *reinterpret_cast<const float4*>(gout + r * ld + col0);
    const float4 y4 = ldg4(x + r * ld + col0);
    float u[4] = {g4.x, g4.y, g4.z, g4.w};
    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (relu && !(fmaf(a4[j], yv[j], c4[j]) > 0.f)) u[j] = 0.f;
      const float d = (yv[j] - k[6][j]) - k[7][j];
      float acc = fmaf(k[3][j], d, k[5][j]);
      acc = fmaf(k[1][j], u[j], acc);
      acc = fmaf(k[2][j], d, acc + k[4][j]);
      o[j] = fmaf(k[0][j], u[j], acc);
    }
    *reinterpret_cast<float4*>(dz + r * ld + col0) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

static int g_bn_small = 1;
extern "C" int sb_set_small_bn(int32_t enable) {
  const int old = g_bn_small;
  g_bn_small = enable ? 1 : 0;
  return old;
}
static bool bn_small_ok(const void* x, int64_t ld, int64_t M, int32_t G) {
  return g_bn_small && G == 1 && M >= 1 && M <= BN_SMALL_MAX_ROWS && (ld % 4 == 0) && ((uintptr_t)x % 16 == 0);
}

// out = act(BN(x)) (+ res): column statistics (training) -> a, c, mean_rstd (+ running buffers) -> apply.
// stats: fp64 [G,2,C] scratch (only touched by the streaming path; zeroed here).
extern "C" int sb_bn_act_fwd(const float* x, int64_t ld, int64_t M, int32_t G, int32_t C, const float* gamma,
                             const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                             int32_t training, int32_t relu, const float* res, float* out, double* stats, float* a, float* c,
                             double* mean_rstd, void* stream) {
  SB_CHECK_ARG(x && out && a && c && G >= 1 && C >= 1 && ld >= C, "sb_bn_act_fwd: bad arguments");
  SB_CHECK_ARG(training ? (M >= 1) : (running_mean && running_var), "sb_bn_act_fwd: eval needs running statistics");
  cudaStream_t st = (cudaStream_t)stream;
  if (bn_small_ok(x, ld, M, G) && ((uintptr_t)out % 16 == 0) && (!res || (uintptr_t)res % 16 == 0)) {
    bn_act_small_fwd_kernel<<<(unsigned)(ld / 4), EW_THREADS, 0, st>>>(x, ld, M, C, gamma, beta, running_mean, running_var,
                                                                       momentum, eps, training, relu, res, out, a, c,
                                                                       mean_rstd);
    SB_CHECK_LAUNCH("sb_bn_act_fwd(small)");
    return SB_OK;
  }
  if (training) {
    SB_CHECK_ARG(stats != nullptr, "sb_bn_act_fwd: stats scratch required");
    SB_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (size_t)G * 2 * C, st));
    const int rc = sb_col_stats(x, ld, M, G, C, stats, stream);
    if (rc) return rc;
  }
  return sb_bn_apply_fwd(x, training ? stats : nullptr, M, G, C, gamma, beta, running_mean, running_var, momentum, eps,
                         training, relu, res, out, ld, M, a, c, mean_rstd, stream);
}

// dz (may alias gout) <- d/dx of act(BN(x)) given gout; dgamma, dbeta.  stats fp64 [G,2,C] and coef fp64 [3,G,C]: scratch of
// the streaming path.
extern "C" int sb_bn_act_bwd(const float* gout, const float* x, const float* a, const float* c, const double* mean_rstd,
                             const float* gamma, int64_t ld, int64_t M, int32_t G, int32_t C, int32_t relu,
                             int32_t training, float* dz, float* dgamma, float* dbeta, double* stats, double* coef,
                             void* stream) {
  SB_CHECK_ARG(gout && x && a && c && mean_rstd && dz && G >= 1 && C >= 1, "sb_bn_act_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (bn_small_ok(x, ld, M, G) && ((uintptr_t)gout % 16 == 0) && ((uintptr_t)dz % 16 == 0)) {
    bn_act_small_bwd_kernel<<<(unsigned)(ld / 4), EW_THREADS, 0, st>>>(gout, x, a, c, mean_rstd, gamma, ld, M, C, relu,
                                                                       training, dz, dgamma, dbeta);
    SB_CHECK_LAUNCH("sb_bn_act_bwd(small)");
    return SB_OK;
  }
  SB_CHECK_ARG(stats && coef, "sb_bn_act_bwd: scratch required");
  SB_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (size_t)G * 2 * C, st));
  int rc = sb_bn_bwd_reduce(gout, x, a, c, mean_rstd, nullptr, ld, M, G, C, relu, stats, stream);
  if (rc) return rc;
  return sb_bn_apply_bwd(gout, x, stats, mean_rstd, relu ? a : nullptr, relu ? c : nullptr, gamma, M, training, dz, dgamma,
                         dbeta, ld, M, G, C, stream);
}

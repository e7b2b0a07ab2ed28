// K2 — the dense per-row contractions of the path (the two Linears inside every GIN layer, rho's MLP, the predictor's
// MLPs) as weight-stationary, persistent fp32 kernels with the neighbouring element-wise work fused in:
//   prologue  (on load of the activation tile): none | BatchNorm-affine | BatchNorm-affine + ReLU
//   epilogue  (on the accumulator tile)       : + bias | ReLU | per-(group, channel) sum / sum-of-squares for the
//                                               BatchNorm that follows (no second pass over the activations)
// Replaces nn.Linear + the `x[~mask] = 0` / `x[mask] = bn(x[mask])` bookkeeping of MaskedMLP
// (Alchemy/sign_net/model_utils/masked_layers.py:54-64) and layers/mlp.py:37-56: on the ragged slot-row layout every
// row is valid, so the masks disappear.
//
// fp32 FFMA on purpose: BASELINE.json's parity bar is 1e-5 relative in fp32, which a single-pass TF32/bf16 tensor-core
// contraction does not meet.  The weight matrix (<= 128 x 128) stays resident in shared memory for the lifetime of a
// persistent CTA; activation tiles are double-buffered through registers so the prologue transform is applied once
// per element.
#include "common.cuh"
#include "../../include/signnet_b200.h"

#include <stdlib.h>

#define LIN_BM 128
#define LIN_BK 32
#define LIN_BKP 36
#define LIN_THREADS 256
#define LIN_MAXG 2

struct LinArgs {
  const float* x;
  long long ldx;
  const float* w;
  long long w_rs, w_cs;
  const float* bias;
  float* y;
  long long ldy;
  long long R;
  int G, K, N, KP;
  int pro;
  const float* pa;
  const float* pc;
  int relu;
  double* stats;
  int accumulate;
  int xvec, yvec;  // 128-bit access allowed on x / y
  int ycols;       // columns of y this launch may write (>= N; the tail N..ycols-1 is zero-filled)
};

template <int BN>
__global__ void __launch_bounds__(LIN_THREADS, (BN == 64) ? 2 : 1) linear_fwd_kernel(const LinArgs a) {
  constexpr int NT = BN / 16;   // output columns per thread (4 or 8)
  constexpr int NV = NT / 4;    // float4 column groups per thread
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                                  // [KP][BN]
  float* As = Ws + (size_t)a.KP * BN;                // [2][LIN_BM][LIN_BKP]
  double* sacc = reinterpret_cast<double*>(As + 2 * LIN_BM * LIN_BKP);  // [LIN_MAXG][2][BN]

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int K = a.K, N = a.N, KP = a.KP;

  // ---- stage the weight: Ws[k][n] = W[n, k], zero padded
  for (int idx = tid; idx < KP * BN; idx += LIN_THREADS) {
    int k, n;
    if (a.w_cs == 1) { n = idx / KP; k = idx - n * KP; } else { k = idx / BN; n = idx - k * BN; }
    float v = 0.f;
    if (k < K && n < N) v = __ldg(a.w + (long long)n * a.w_rs + (long long)k * a.w_cs);
    Ws[k * BN + n] = v;
  }
  if (a.stats)
    for (int idx = tid; idx < LIN_MAXG * 2 * BN; idx += LIN_THREADS) sacc[idx] = 0.0;

  const long long tpg = (a.R + LIN_BM - 1) / LIN_BM;  // tiles per group
  const long long ntiles = tpg * a.G;
  const int nchunks = KP / LIN_BK;

  float4 pre[4];
  auto load_chunk = [&](long long tile, int chunk) {
    const int g = (int)(tile / tpg);
    const long long row0 = (tile - (long long)g * tpg) * LIN_BM;
    const int rows = (int)((a.R - row0 < LIN_BM) ? (a.R - row0) : LIN_BM);
    const long long base = (long long)g * a.R + row0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int f = tid + LIN_THREADS * q;
      const int row = f >> 3, c4 = f & 7;
      const int col = chunk * LIN_BK + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows && col < K) {
        const float* p = a.x + (base + row) * a.ldx + col;
        if (a.xvec) {
          v = ldg4(p);
        } else {
          v.x = __ldg(p);
          if (col + 1 < K) v.y = __ldg(p + 1);
          if (col + 2 < K) v.z = __ldg(p + 2);
          if (col + 3 < K) v.w = __ldg(p + 3);
        }
        if (a.pro) {
          const float* pa = a.pa + (long long)g * K + col;
          const float* pc = a.pc + (long long)g * K + col;
          float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (col + j < K) {
              float u = fmaf(__ldg(pa + j), t[j], __ldg(pc + j));
              t[j] = (a.pro == 2) ? fmaxf(u, 0.f) : u;
            } else {
              t[j] = 0.f;
            }
          }
          v = make_float4(t[0], t[1], t[2], t[3]);
        } else {
          if (col + 1 >= K) v.y = 0.f;
          if (col + 2 >= K) v.z = 0.f;
          if (col + 3 >= K) v.w = 0.f;
        }
      }
      pre[q] = v;
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int f = tid + LIN_THREADS * q;
      const int row = f >> 3, c4 = f & 7;
      *reinterpret_cast<float4*>(As + ((size_t)buf * LIN_BM + row) * LIN_BKP + c4 * 4) = pre[q];
    }
  };

  float acc[8][NT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;

  long long tile = blockIdx.x;
  int chunk = 0, buf = 0;
  if (tile < ntiles) {
    load_chunk(tile, 0);
    store_chunk(0);
  }
  __syncthreads();

  while (tile < ntiles) {
    long long ntile = tile;
    int nchunk = chunk + 1;
    if (nchunk == nchunks) { nchunk = 0; ntile = tile + gridDim.x; }
    const bool has_next = ntile < ntiles;
    if (has_next) load_chunk(ntile, nchunk);

    const float* Ab = As + (size_t)buf * LIN_BM * LIN_BKP;
    const float* Wb = Ws + (size_t)chunk * LIN_BK * BN;
#pragma unroll
    for (int kk = 0; kk < LIN_BK; kk += 4) {
      float4 av[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = *reinterpret_cast<const float4*>(Ab + (ty + 16 * i) * LIN_BKP + kk);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        float wv[NT];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float4 w4 = *reinterpret_cast<const float4*>(Wb + (kk + k4) * BN + v * 64 + tx * 4);
          wv[v * 4 + 0] = w4.x; wv[v * 4 + 1] = w4.y; wv[v * 4 + 2] = w4.z; wv[v * 4 + 3] = w4.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float aval = (k4 == 0) ? av[i].x : (k4 == 1) ? av[i].y : (k4 == 2) ? av[i].z : av[i].w;
#pragma unroll
          for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(aval, wv[j], acc[i][j]);
        }
      }
    }

    if (has_next) store_chunk(buf ^ 1);

    if (chunk == nchunks - 1) {
      // ---------------------------------------------------------------- epilogue for this row tile
      const int g = (int)(tile / tpg);
      const long long row0 = (tile - (long long)g * tpg) * LIN_BM;
      const int rows = (int)((a.R - row0 < LIN_BM) ? (a.R - row0) : LIN_BM);
      const long long base = (long long)g * a.R + row0;
      double ssum[NT], ssq[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) ssum[j] = 0.0, ssq[j] = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = ty + 16 * i;
        if (row < rows) {
          float* yrow = a.y + (base + row) * a.ldy;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int col0 = v * 64 + tx * 4;
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = col0 + j;
              float t = acc[i][v * 4 + j];
              if (col < N) {
                if (a.bias) t += __ldg(a.bias + col);
                if (a.accumulate) t += yrow[col];
                if (a.relu) t = fmaxf(t, 0.f);
                if (a.stats) {
                  ssum[v * 4 + j] += (double)t;
                  ssq[v * 4 + j] += (double)t * (double)t;
                }
              } else {
                t = 0.f;
              }
              o[j] = t;
            }
            if (a.yvec) {
              if (col0 < a.ycols) *reinterpret_cast<float4*>(yrow + col0) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col0 + j < a.ycols) yrow[col0 + j] = o[j];
            }
          }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
      }
      if (a.stats) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 16);
          ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], 16);
        }
        if ((tid & 31) < 16) {
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = v * 64 + tx * 4 + j;
              if (col < N) {
                atomicAdd(&sacc[(g * 2 + 0) * BN + col], ssum[v * 4 + j]);
                atomicAdd(&sacc[(g * 2 + 1) * BN + col], ssq[v * 4 + j]);
              }
            }
        }
      }
    }
    __syncthreads();
    buf ^= 1;
    tile = ntile;
    chunk = nchunk;
  }

  if (a.stats) {
    __syncthreads();
    for (int idx = tid; idx < a.G * 2 * BN; idx += LIN_THREADS) {
      const int col = idx % BN, gj = idx / BN;
      if (col < N && sacc[idx] != 0.0) atomicAdd(a.stats + (long long)gj * N + col, sacc[idx]);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Small-row variant (same contract): the [N, d] contractions of the predictor / rho / the DGL MLPs have only a few
// thousand rows; with 128-row tiles they fill 23 of the 148 SMs and every CTA first stages the whole weight
// (51 us per launch at 2 922 rows, ncu launch list profiles/r2e_launches_b128.csv - 18 % of a 128-graph step).  Here a
// CTA owns 32 rows x <= 128 columns and streams BOTH operands K-chunk by K-chunk through shared memory (the weight is
// 64 KB and L2-resident), so a few-thousand-row problem spreads over ~100 CTAs and nothing waits for a 64 KB prologue.
// Thread (tx = lane, ty = warp): rows 4 ty .. 4 ty + 3, columns tx, tx + 32, tx + 64, tx + 96 (conflict-free LDS, the
// activation value is a warp-wide broadcast, every global store of a warp is one 128-byte line).
#define LS_BM 32
#define LS_BK 32
#define LS_WP 129   // padded row of the weight chunk [LS_BK][128]

__global__ void __launch_bounds__(256) linear_small_kernel(const LinArgs a) {
  __shared__ float Xs[2][LS_BM][LS_BK + 1];
  __shared__ float Ws[2][LS_BK][LS_WP];
  __shared__ double sacc[2][128];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int K = a.K, N = a.N;
  const long long tpg = (a.R + LS_BM - 1) / LS_BM;
  const long long tile = blockIdx.x;
  const int g = (int)(tile / tpg);
  const long long row0 = (tile - (long long)g * tpg) * LS_BM;
  const int rows = (int)((a.R - row0 < LS_BM) ? (a.R - row0) : LS_BM);
  const long long base = (long long)g * a.R + row0;
  if (a.stats) sacc[tid >> 7][tid & 127] = 0.0;

  // Register-staged, double-buffered K-chunks: the loads of chunk c + 1 are in flight while chunk c is multiplied (the
  // first version waited for every chunk's round trip: 27 us per launch at 2 922 rows, 4 chunks).
  float xr[4], wr[16];
  const bool w_kmajor = (a.w_cs == 1);
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + 256 * q, r = idx >> 5, k = k0 + (idx & 31);
      float v = 0.f;
      if (r < rows && k < K) {
        v = __ldg(a.x + (base + r) * a.ldx + k);
        if (a.pro) {
          v = fmaf(__ldg(a.pa + (long long)g * K + k), v, __ldg(a.pc + (long long)g * K + k));
          if (a.pro == 2) v = fmaxf(v, 0.f);
        }
      }
      xr[q] = v;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int idx = tid + 256 * q;
      int n, kk;
      if (w_kmajor) { kk = idx & 31; n = idx >> 5; } else { n = idx & 127; kk = idx >> 7; }
      const int k = k0 + kk;
      wr[q] = (n < N && k < K) ? __ldg(a.w + (long long)n * a.w_rs + (long long)k * a.w_cs) : 0.f;
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + 256 * q;
      Xs[buf][idx >> 5][idx & 31] = xr[q];
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int idx = tid + 256 * q;
      int n, kk;
      if (w_kmajor) { kk = idx & 31; n = idx >> 5; } else { n = idx & 127; kk = idx >> 7; }
      Ws[buf][kk][n] = wr[q];
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += LS_BK, buf ^= 1) {
    const bool more = k0 + LS_BK < K;
    if (more) load_chunk(k0 + LS_BK);            // in flight during the multiply below
#pragma unroll 8
    for (int kk = 0; kk < LS_BK; ++kk) {
      float xv[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = Xs[buf][ty * 4 + i][kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = Ws[buf][kk][tx + 32 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
    }
    if (more) store_chunk(buf ^ 1);              // the other buffer: last read two iterations ago (barrier below)
    __syncthreads();
  }

  double ssum[4] = {0.0, 0.0, 0.0, 0.0}, ssq[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (r >= rows) continue;
    float* yrow = a.y + (base + r) * a.ldy;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = tx + 32 * j;
      float t = acc[i][j];
      if (col < N) {
        if (a.bias) t += __ldg(a.bias + col);
        if (a.accumulate) t += yrow[col];
        if (a.relu) t = fmaxf(t, 0.f);
        ssum[j] += (double)t;
        ssq[j] += (double)t * (double)t;
        yrow[col] = t;
      } else if (col < a.ycols) {
        yrow[col] = 0.f;
      }
    }
  }
  if (a.stats) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = tx + 32 * j;
      if (col < N) {
        atomicAdd(&sacc[0][col], ssum[j]);
        atomicAdd(&sacc[1][col], ssq[j]);
      }
    }
    __syncthreads();
    const int col = tid & 127, which = tid >> 7;
    if (col < N && sacc[which][col] != 0.0)
      atomicAdd(a.stats + (long long)(g * 2 + which) * N + col, sacc[which][col]);
  }
}

// rows (all groups) at or below which the small-row kernel is used instead of the tcgen05 / 128-row FFMA kernels
#define LS_MAX_ROWS 8192

static int launch_linear_small(const LinArgs& a, cudaStream_t st) {
  const long long ntiles = sb_ceil_div(a.R, LS_BM) * a.G;
  linear_small_kernel<<<(unsigned)ntiles, 256, 0, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_fwd(small rows)");
  return SB_OK;
}

static size_t lin_smem_bytes(int KP, int BN) {
  return ((size_t)KP * BN + 2 * LIN_BM * LIN_BKP) * sizeof(float) + (size_t)LIN_MAXG * 2 * BN * sizeof(double);
}

// tensor-core path (linear_tc.cu)
int sb_linear_tc_launch(const float* x, int64_t ldx, const float* w, int64_t w_rs, int64_t w_cs, const float* bias,
                        float* y, int64_t ldy, int64_t R, int32_t G, int32_t K, int32_t N, int32_t pro,
                        const float* pa, const float* pc, int32_t relu, double* stats, int32_t accumulate,
                        int32_t ycols, cudaStream_t st);
// -1 undecided, 0 FFMA, 1 tcgen05 (default: Linear = linear_tc.cu; weight gradient = the TMA-fed wgrad_tc_tma.cu for
// N = K = 128, the register-fed wgrad_tc.cu otherwise), 2 = tcgen05 with the register-fed weight gradient everywhere
// (kept so that tests can compare the two weight-gradient kernels bit for bit).  Round-2 measurements that decided this
// (profiles/r2a_pair_check_mode*.log): the TMA-fed weight gradient is bit-identical and 17 % faster (278 vs 338 us at the
// phi size); the CTA-pair, TMA-fed and weight-in-tensor-memory Linear variants were all slower than linear_tc.cu
// (393 / 310 / 314 vs 292 us) and have been removed.  So was a 256-row-tile variant (bit-identical, but 353 us: with one
// 256-column accumulator and a 3-stage ring the load -> split -> MMA -> drain phases of a tile serialise,
// profiles/r2i_linear_tc256.log) and a second MMA-issuing warp in linear_tc.cu (no change: that kernel is bound by the
// hand-offs of its two-stage ring; the fused kernel and the weight gradient did gain from it).
static int g_use_tc = -1;
static int g_small_rows = 1;   // SB_LINEAR_SMALL=0 / sb_set_small_rows(0): problems with few rows take the big-tile kernels
static int g_last_variant = -1, g_last_wgrad_variant = -1;
extern "C" int sb_set_small_rows(int32_t enable) {
  const int old = g_small_rows;
  g_small_rows = enable ? 1 : 0;
  return old;
}
extern "C" int sb_last_linear_kernel(void) { return g_last_variant; }
extern "C" int sb_last_wgrad_kernel(void) { return g_last_wgrad_variant; }
extern "C" int sb_set_tensor_cores(int32_t enable) {
  const int old = g_use_tc;
  g_use_tc = (enable == 2) ? 2 : (enable ? 1 : 0);
  return old;
}
static bool use_tc() {
  if (g_use_tc < 0) {
    const char* e = getenv("SB_DISABLE_TC");
    g_use_tc = (e && e[0] == '1') ? 0 : 1;
  }
  return g_use_tc >= 1;
}

int sb_rank1_fwd_launch(const float* x, int64_t ldx, const float* w, int64_t w_rs, const float* bias, float* y,
                        int64_t ldy, int64_t R, int32_t G, int32_t N, int32_t ycols, int32_t pro, const float* pa,
                        const float* pc, int32_t relu, double* stats, cudaStream_t st);
int sb_rowdot_fwd_launch(const float* x, int64_t ldx, const float* w, int64_t w_cs, const float* bias, float* y,
                         int64_t ldy, int64_t R, int32_t G, int32_t K, int32_t pro, const float* pa, const float* pc,
                         int32_t relu, cudaStream_t st);
int sb_rank1_wgrad_launch(const float* gy, int64_t ldg, const float* x, int64_t ldx, int64_t R, int32_t G, int32_t N,
                          int32_t pro, const float* pa, const float* pc, float* dw, int64_t dw_rs, float* db,
                          int32_t accumulate, float* workspace, cudaStream_t st);

template <int BN>
static int launch_linear(const LinArgs& a, cudaStream_t st) {
  static size_t configured = 0;
  const size_t smem = lin_smem_bytes(a.KP, BN);
  if (smem > configured) {
    SB_CUDA(cudaFuncSetAttribute(linear_fwd_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)lin_smem_bytes(128, BN)));
    configured = lin_smem_bytes(128, BN);
  }
  const long long ntiles = sb_ceil_div(a.R, LIN_BM) * a.G;
  long long grid = (long long)sb_num_sms() * ((BN == 64) ? 2 : 1);
  if (grid > ntiles) grid = ntiles;
  linear_fwd_kernel<BN><<<(unsigned)grid, LIN_THREADS, smem, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_fwd");
  return SB_OK;
}

// y[g*R + r, n] (+)= sum_k f(x[g*R + r, k]) * W[n, k] + bias[n]   for n < N; columns N..ldy-1 of y are zero-filled.
extern "C" int sb_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t w_rs, int64_t w_cs,
                             const float* bias, float* y, int64_t ldy, int64_t R, int32_t G, int32_t K, int32_t N,
                             int32_t pro, const float* pa, const float* pc, int32_t relu, double* stats,
                             int32_t accumulate, void* stream) {
  SB_CHECK_ARG(R >= 0 && G >= 1 && G <= LIN_MAXG && K >= 1 && N >= 1, "sb_linear_fwd: bad sizes R=%lld G=%d K=%d N=%d",
               (long long)R, G, K, N);
  SB_CHECK_ARG(ldx >= K && ldy >= N, "sb_linear_fwd: leading dims too small");
  SB_CHECK_ARG(pro >= 0 && pro <= 2 && (pro == 0 || (pa && pc)), "sb_linear_fwd: bad prologue");
  if (R == 0) return SB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // degenerate shapes of the first phi layer: outer product (K = 1) / row dot product (N = 1) streaming kernels
  if (K == 1 && N <= 128 && !accumulate) {
    const int rc = sb_rank1_fwd_launch(x, ldx, w, w_rs, bias, y, ldy, R, G, N, (int)(ldy < 128 ? ldy : 128), pro, pa, pc,
                                       relu, stats, st);
    if (rc != SB_ERR_UNSUPPORTED) return rc;
  }
  if (N == 1 && !accumulate && !stats && (!pro || K <= 128)) {
    const int rc = sb_rowdot_fwd_launch(x, ldx, w, w_cs, bias, y, ldy, R, G, K, pro, pa, pc, relu, st);
    if (rc != SB_ERR_UNSUPPORTED) return rc;
  }
  // Host-side blocking over N (<= 128 per launch) and K (<= 128 per launch, accumulating).
  for (int n0 = 0; n0 < N; n0 += 128) {
    const int nn = (N - n0 < 128) ? (N - n0) : 128;
    for (int k0 = 0; k0 < K; k0 += 128) {
      const int kk = (K - k0 < 128) ? (K - k0) : 128;
      const bool last_k = (k0 + 128 >= K);
      LinArgs a;
      a.x = x + k0; a.ldx = ldx;
      a.w = w + (long long)n0 * w_rs + (long long)k0 * w_cs; a.w_rs = w_rs; a.w_cs = w_cs;
      a.bias = (last_k && bias) ? bias + n0 : nullptr;
      a.y = y + n0;
      a.ldy = ldy;
      a.R = R; a.G = G; a.K = kk; a.N = nn;
      a.KP = (int)sb_ceil_div(kk, LIN_BK) * LIN_BK;
      a.pro = pro; a.pa = pa ? pa + k0 : nullptr; a.pc = pc ? pc + k0 : nullptr;
      a.relu = last_k ? relu : 0;
      a.stats = last_k ? (stats ? stats + n0 : nullptr) : nullptr;
      a.accumulate = (k0 > 0) ? 1 : accumulate;
      SB_CHECK_ARG(!(pro && K > 128), "sb_linear_fwd: prologue with K > 128 unsupported");
      SB_CHECK_ARG(!(stats && N > 128), "sb_linear_fwd: stats with N > 128 unsupported");
      a.xvec = (ldx % 4 == 0) && ((uintptr_t)a.x % 16 == 0);
      a.yvec = (ldy % 4 == 0) && ((uintptr_t)a.y % 16 == 0);
      const bool last_n = (n0 + 128 >= N);
      a.ycols = last_n ? (int)((ldy - n0 < 128) ? (ldy - n0) : 128) : 128;
      int rc = SB_ERR_UNSUPPORTED;
      int variant = 0;
      if (g_small_rows && a.R * a.G <= LS_MAX_ROWS) {
        rc = launch_linear_small(a, st);
        if (rc != SB_OK) return rc;
        g_last_variant = 4;
        continue;
      }
      if (use_tc()) {
        variant = 1;
        rc = sb_linear_tc_launch(a.x, a.ldx, a.w, a.w_rs, a.w_cs, a.bias, a.y, a.ldy, a.R, a.G, a.K, a.N, a.pro, a.pa,
                                 a.pc, a.relu, a.stats, a.accumulate, a.ycols, st);
      }
      if (rc == SB_ERR_UNSUPPORTED) variant = 0;
      g_last_variant = variant;
      if (rc == SB_ERR_UNSUPPORTED)
        rc = (nn <= 64 && a.ycols <= 64) ? launch_linear<64>(a, st) : launch_linear<128>(a, st);
      if (rc != SB_OK) return rc;
    }
  }
  return SB_OK;
}

// =====================================================================================================================
// Weight gradient:  dW[n, k] = sum_rows g[row, n] * f(x[row, k]),  dbias[n] = sum_rows g[row, n]
// (f = the same prologue as the forward, recomputed instead of stored).  Tall-skinny reduction: persistent CTAs each
// own a strided subset of 32-row chunks, keep a full [BN x BK] partial in registers, write it once to a workspace and
// a second tiny kernel adds the per-CTA partials in a fixed order (deterministic, no float atomics).
// =====================================================================================================================
#define WG_BR 32

struct WgArgs {
  const float* g;
  long long ldg;
  const float* x;
  long long ldx;
  long long R;
  int G, N, K;
  int pro;
  const float* pa;
  const float* pc;
  float* part_w;   // [grid][BN][BK]
  float* part_b;   // [grid][BN] or null
  int gvec, xvec;
};

template <int BN, int BK>
__global__ void __launch_bounds__(LIN_THREADS, 1) linear_wgrad_kernel(const WgArgs a) {
  constexpr int MI = BN / 16;          // dW rows (n) per thread
  constexpr int NT = BK / 16;          // dW cols (k) per thread
  constexpr int NV = NT / 4;
  constexpr int GQ = BN / 32;          // float4 loads of g per thread per chunk
  constexpr int XQ = BK / 32;          // float4 loads of x per thread per chunk
  extern __shared__ __align__(16) float smem[];
  float* Gt = smem;                                   // [2][BN][LIN_BKP]   (transposed: [n][row])
  float* Xs = Gt + 2 * BN * LIN_BKP;                  // [2][WG_BR][BK]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, wrp = tid >> 5;
  const int N = a.N, K = a.K;

  const long long cpg = (a.R + WG_BR - 1) / WG_BR;
  const long long nchunks = cpg * a.G;

  float4 pg[GQ], px[XQ];
  float dbs[GQ][4];
#pragma unroll
  for (int q = 0; q < GQ; ++q) dbs[q][0] = dbs[q][1] = dbs[q][2] = dbs[q][3] = 0.f;

  auto load_chunk = [&](long long c) {
    const int g = (int)(c / cpg);
    const long long row0 = (c - (long long)g * cpg) * WG_BR;
    const int rows = (int)((a.R - row0 < WG_BR) ? (a.R - row0) : WG_BR);
    const long long base = (long long)g * a.R + row0;
#pragma unroll
    for (int q = 0; q < GQ; ++q) {  // g: lane -> row (transposed smem store is then conflict-free)
      const int row = lane, c4 = wrp + 8 * q;
      const int col = c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows && col < N) {
        const float* p = a.g + (base + row) * a.ldg + col;
        if (a.gvec) {
          v = ldg4(p);
        } else {
          v.x = __ldg(p);
          if (col + 1 < N) v.y = __ldg(p + 1);
          if (col + 2 < N) v.z = __ldg(p + 2);
          if (col + 3 < N) v.w = __ldg(p + 3);
        }
        if (col + 1 >= N) v.y = 0.f;
        if (col + 2 >= N) v.z = 0.f;
        if (col + 3 >= N) v.w = 0.f;
      }
      pg[q] = v;
    }
#pragma unroll
    for (int q = 0; q < XQ; ++q) {
      const int f = tid + LIN_THREADS * q;
      const int row = f / (BK / 4), c4 = f % (BK / 4);
      const int col = c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows && col < K) {
        const float* p = a.x + (base + row) * a.ldx + col;
        if (a.xvec) {
          v = ldg4(p);
        } else {
          v.x = __ldg(p);
          if (col + 1 < K) v.y = __ldg(p + 1);
          if (col + 2 < K) v.z = __ldg(p + 2);
          if (col + 3 < K) v.w = __ldg(p + 3);
        }
        float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (col + j < K) {
            if (a.pro) {
              float u = fmaf(__ldg(a.pa + (long long)g * K + col + j), t[j], __ldg(a.pc + (long long)g * K + col + j));
              t[j] = (a.pro == 2) ? fmaxf(u, 0.f) : u;
            }
          } else {
            t[j] = 0.f;
          }
        }
        v = make_float4(t[0], t[1], t[2], t[3]);
      }
      px[q] = v;
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int q = 0; q < GQ; ++q) {
      const int row = lane, c4 = wrp + 8 * q;
      float* dst = Gt + ((size_t)buf * BN + c4 * 4) * LIN_BKP + row;
      dst[0] = pg[q].x; dst[LIN_BKP] = pg[q].y; dst[2 * LIN_BKP] = pg[q].z; dst[3 * LIN_BKP] = pg[q].w;
      dbs[q][0] += pg[q].x; dbs[q][1] += pg[q].y; dbs[q][2] += pg[q].z; dbs[q][3] += pg[q].w;
    }
#pragma unroll
    for (int q = 0; q < XQ; ++q) {
      const int f = tid + LIN_THREADS * q;
      const int row = f / (BK / 4), c4 = f % (BK / 4);
      *reinterpret_cast<float4*>(Xs + ((size_t)buf * WG_BR + row) * BK + c4 * 4) = px[q];
    }
  };

  float acc[MI][NT];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;

  long long c = blockIdx.x;
  int buf = 0;
  if (c < nchunks) {
    load_chunk(c);
    store_chunk(0);
  }
  __syncthreads();
  while (c < nchunks) {
    const long long nc = c + gridDim.x;
    const bool has_next = nc < nchunks;
    if (has_next) load_chunk(nc);
    const float* Gb = Gt + (size_t)buf * BN * LIN_BKP;
    const float* Xb = Xs + (size_t)buf * WG_BR * BK;
#pragma unroll
    for (int rr = 0; rr < WG_BR; rr += 4) {
      float4 gv[MI];
#pragma unroll
      for (int i = 0; i < MI; ++i) gv[i] = *reinterpret_cast<const float4*>(Gb + (ty + 16 * i) * LIN_BKP + rr);
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        float xv[NT];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float4 x4 = *reinterpret_cast<const float4*>(Xb + (rr + r4) * BK + v * 64 + tx * 4);
          xv[v * 4 + 0] = x4.x; xv[v * 4 + 1] = x4.y; xv[v * 4 + 2] = x4.z; xv[v * 4 + 3] = x4.w;
        }
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const float gval = (r4 == 0) ? gv[i].x : (r4 == 1) ? gv[i].y : (r4 == 2) ? gv[i].z : gv[i].w;
#pragma unroll
          for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(gval, xv[j], acc[i][j]);
        }
      }
    }
    if (has_next) store_chunk(buf ^ 1);
    __syncthreads();
    buf ^= 1;
    c = nc;
  }

  float* pw = a.part_w + (size_t)blockIdx.x * BN * BK;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int v = 0; v < NV; ++v)
      *reinterpret_cast<float4*>(pw + (size_t)(ty + 16 * i) * BK + v * 64 + tx * 4) =
          make_float4(acc[i][v * 4 + 0], acc[i][v * 4 + 1], acc[i][v * 4 + 2], acc[i][v * 4 + 3]);
  if (a.part_b) {
#pragma unroll
    for (int q = 0; q < GQ; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = warp_sum(dbs[q][j]);
        if (lane == 0) a.part_b[(size_t)blockIdx.x * BN + (wrp + 8 * q) * 4 + j] = t;
      }
  }
}

// Four lanes per output element: lane q adds the partials p = q, q+4, ... (two independent chains each), then a fixed
// two-step shuffle tree - the same order on every run.  (One thread per element walked 148 dependent loads: 20 us per
// call, 37 calls per step.)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b, int nparts,
                                    int BN, int BK, int N, int K, float* __restrict__ dw, long long rs, long long cs,
                                    float* __restrict__ db, int accumulate) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = t >> 2, q = t & 3;
  const int total = N * K + (db ? N : 0);
  const bool live = idx < total;
  const float* src = nullptr;
  size_t pstride = 0;
  if (live) {
    if (idx < N * K) {
      const int n = idx / K, k = idx - n * K;
      src = part_w + (size_t)n * BK + k;
      pstride = (size_t)BN * BK;
    } else {
      src = part_b + (idx - N * K);
      pstride = (size_t)BN;
    }
  }
  float s0 = 0.f, s1 = 0.f;
  if (live) {
    int p = q;
    for (; p + 4 < nparts; p += 8) {
      s0 += __ldg(src + (size_t)p * pstride);
      s1 += __ldg(src + (size_t)(p + 4) * pstride);
    }
    if (p < nparts) s0 += __ldg(src + (size_t)p * pstride);
  }
  float s = s0 + s1;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (live && q == 0) {
    if (idx < N * K) {
      const int n = idx / K, k = idx - n * K;
      float* dst = dw + n * rs + k * cs;
      *dst = accumulate ? (*dst + s) : s;
    } else {
      const int n = idx - N * K;
      db[n] = accumulate ? (db[n] + s) : s;
    }
  }
}

int sb_wgrad_reduce_launch(const float* part_w, const float* part_b, int nparts, int BN, int BK, int N, int K, float* dw,
                           long long rs, long long cs, float* db, int accumulate, cudaStream_t st) {
  const int total = N * K + (db ? N : 0);
  // four lanes per output element
  wgrad_reduce_kernel<<<(unsigned)sb_ceil_div((long long)total * 4, 128), 128, 0, st>>>(part_w, part_b, nparts, BN, BK, N,
                                                                                       K, dw, rs, cs, db, accumulate);
  SB_CHECK_LAUNCH("sb_linear_wgrad(reduce)");
  return SB_OK;
}
int sb_wgrad_tc_launch(const float* gy, int64_t ldg, const float* x, int64_t ldx, int64_t R, int32_t G, int32_t N,
                       int32_t K, int32_t pro, const float* pa, const float* pc, float* dw, int64_t dw_rs,
                       int64_t dw_cs, float* db, int32_t accumulate, float* workspace, cudaStream_t st);

int sb_wgrad_tc_tma_launch(const float* gy, int64_t ldg, const float* x, int64_t ldx, int64_t R, int32_t G, int32_t N,
                           int32_t K, int32_t pro, const float* pa, const float* pc, float* dw, int64_t dw_rs,
                           int64_t dw_cs, float* db, int32_t accumulate, float* workspace, cudaStream_t st);

template <int BN, int BK>
static int launch_wgrad(WgArgs a, int N, int K, float* dw, long long rs, long long cs, float* db, int accumulate,
                        float* workspace, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = ((size_t)2 * BN * LIN_BKP + 2 * WG_BR * BK) * sizeof(float);
  if (!configured) {
    SB_CUDA(cudaFuncSetAttribute(linear_wgrad_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long nchunks = sb_ceil_div(a.R, WG_BR) * a.G;
  long long grid = sb_num_sms();
  if (grid > nchunks) grid = nchunks;
  a.part_w = workspace;
  a.part_b = db ? workspace + (size_t)grid * BN * BK : nullptr;
  linear_wgrad_kernel<BN, BK><<<(unsigned)grid, LIN_THREADS, smem, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_wgrad");
  return sb_wgrad_reduce_launch(a.part_w, a.part_b, (int)grid, BN, BK, N, K, dw, rs, cs, db, accumulate, st);
}

extern "C" int64_t sb_linear_wgrad_workspace_floats(void) {
  return (int64_t)sb_num_sms() * (128 * 128 + 128);
}

// dw[n*rs + k*cs] (+)= sum_{g,r} gy[g*R + r, n] * f(x[g*R + r, k]);  db[n] (+)= sum gy[., n]
extern "C" int sb_linear_wgrad(const float* gy, int64_t ldg, const float* x, int64_t ldx, int64_t R, int32_t G,
                               int32_t N, int32_t K, int32_t pro, const float* pa, const float* pc, float* dw,
                               int64_t dw_rs, int64_t dw_cs, float* db, int32_t accumulate, float* workspace,
                               void* stream) {
  SB_CHECK_ARG(R >= 0 && G >= 1 && N >= 1 && K >= 1 && ldg >= N && ldx >= K, "sb_linear_wgrad: bad sizes");
  SB_CHECK_ARG(pro >= 0 && pro <= 2 && (pro == 0 || (pa && pc)), "sb_linear_wgrad: bad prologue");
  SB_CHECK_ARG(workspace != nullptr, "sb_linear_wgrad: workspace required");
  cudaStream_t st = (cudaStream_t)stream;
  if (R == 0) {
    if (!accumulate) {
      // empty batch: gradients are zero
      for (int n = 0; n < N; ++n) SB_CUDA(cudaMemsetAsync(dw + n * dw_rs, 0, sizeof(float) * (dw_cs == 1 ? K : 1), st));
      if (db) SB_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, st));
    }
    return SB_OK;
  }
  if (K == 1 && N <= 128) {   // first phi layer: dW0[n] = sum_rows g[row, n] * f(x[row])
    const int rc = sb_rank1_wgrad_launch(gy, ldg, x, ldx, R, G, N, pro, pa, pc, dw, dw_rs, db, accumulate, workspace, st);
    if (rc != SB_ERR_UNSUPPORTED) return rc;
  }
  for (int n0 = 0; n0 < N; n0 += 128) {
    const int nn = (N - n0 < 128) ? (N - n0) : 128;
    for (int k0 = 0; k0 < K; k0 += 128) {
      const int kk = (K - k0 < 128) ? (K - k0) : 128;
      WgArgs a;
      a.g = gy + n0; a.ldg = ldg; a.x = x + k0; a.ldx = ldx; a.R = R; a.G = G; a.N = nn; a.K = kk;
      a.pro = pro; a.pa = pa ? pa + k0 : nullptr; a.pc = pc ? pc + k0 : nullptr;
      SB_CHECK_ARG(!(pro && K > 128), "sb_linear_wgrad: prologue with K > 128 unsupported");
      a.gvec = (ldg % 4 == 0) && ((uintptr_t)a.g % 16 == 0);
      a.xvec = (ldx % 4 == 0) && ((uintptr_t)a.x % 16 == 0);
      a.part_w = nullptr; a.part_b = nullptr;
      float* dwp = dw + (long long)n0 * dw_rs + (long long)k0 * dw_cs;
      float* dbp = (db && k0 == 0) ? db + n0 : nullptr;
      int rc = SB_ERR_UNSUPPORTED;
      if (use_tc()) {
        g_last_wgrad_variant = 3;
        if (g_use_tc == 1)   // TMA-fed operand ring (N == K == 128 only)
          rc = sb_wgrad_tc_tma_launch(a.g, a.ldg, a.x, a.ldx, a.R, a.G, a.N, a.K, a.pro, a.pa, a.pc, dwp, dw_rs, dw_cs,
                                      dbp, accumulate, workspace, st);
        if (rc == SB_ERR_UNSUPPORTED) {
          g_last_wgrad_variant = 1;
          rc = sb_wgrad_tc_launch(a.g, a.ldg, a.x, a.ldx, a.R, a.G, a.N, a.K, a.pro, a.pa, a.pc, dwp, dw_rs, dw_cs, dbp,
                                  accumulate, workspace, st);
        }
      }
      if (rc == SB_ERR_UNSUPPORTED) g_last_wgrad_variant = 0;
      if (rc != SB_ERR_UNSUPPORTED) {
        if (rc != SB_OK) return rc;
        continue;
      }
      if (nn <= 64 && kk <= 64) rc = launch_wgrad<64, 64>(a, nn, kk, dwp, dw_rs, dw_cs, dbp, accumulate, workspace, st);
      else if (nn <= 64) rc = launch_wgrad<64, 128>(a, nn, kk, dwp, dw_rs, dw_cs, dbp, accumulate, workspace, st);
      else if (kk <= 64) rc = launch_wgrad<128, 64>(a, nn, kk, dwp, dw_rs, dw_cs, dbp, accumulate, workspace, st);
      else rc = launch_wgrad<128, 128>(a, nn, kk, dwp, dw_rs, dw_cs, dbp, accumulate, workspace, st);
      if (rc != SB_OK) return rc;
    }
  }
  return SB_OK;
}

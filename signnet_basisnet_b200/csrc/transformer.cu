// K7 — rho of the PyG trees: per-node self-attention over the k_b valid eigenvector slots + LayerNorm.
// Replaces ScaledDotProductAttention / MultiHeadAttention / MaskedLN of
// Alchemy/sign_net/model_utils/transformer_module.py:44-102 and masked_layers.py:22-32.
//
// A "sequence" is the set of slot rows of one node:  token j of node i of graph b is row row_ptr[b] + j*n_b + i.
// Only valid tokens exist in the slot-row layout, so the reference's pairwise mask (fill -1e10 -> softmax -> * mask,
// transformer_module.py:52-56) reduces to a softmax over the k_b real keys: exp(-1e10 - max) is exactly 0 in fp32.
// One CTA per (node, head): Q/K/V slices of <= 37 x d_k floats live in shared memory, scores/probabilities too.
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define ATT_THREADS 128

#include "attention.cuh"

// shared layout helper
struct AttSmem {
  float* Q;   // [kb][dk]   (already divided by the temperature)
  float* K;   // [kb][dk]
  float* V;   // [kb][dk]
  float* P;   // [kb][kb]
};

__device__ __forceinline__ void att_load(const AttArgs& a, long long r0, int n, int kb, int h, AttSmem s) {
  const int dk = a.dk;
  for (int idx = threadIdx.x; idx < kb * dk; idx += blockDim.x) {
    const int j = idx / dk, c = idx - j * dk;
    const long long off = (r0 + (long long)j * n) * a.ld + h * dk + c;
    s.Q[idx] = __fdiv_rn(__ldg(a.q + off), a.inv_temp_div);
    s.K[idx] = __ldg(a.k + off);
    s.V[idx] = __ldg(a.v + off);
  }
}
__device__ __forceinline__ void att_probs(const AttArgs& a, int kb, AttSmem s) {
  const int dk = a.dk;
  for (int idx = threadIdx.x; idx < kb * kb; idx += blockDim.x) {
    const int j1 = idx / kb, j2 = idx - j1 * kb;
    float acc = 0.f;
    for (int c = 0; c < dk; ++c) acc = fmaf(s.Q[j1 * dk + c], s.K[j2 * dk + c], acc);
    s.P[idx] = acc;
  }
  __syncthreads();
  for (int j1 = threadIdx.x; j1 < kb; j1 += blockDim.x) {
    float m = -INFINITY;
    for (int j2 = 0; j2 < kb; ++j2) m = fmaxf(m, s.P[j1 * kb + j2]);
    float sum = 0.f;
    for (int j2 = 0; j2 < kb; ++j2) {
      const float e = expf(s.P[j1 * kb + j2] - m);
      s.P[j1 * kb + j2] = e;
      sum += e;
    }
    for (int j2 = 0; j2 < kb; ++j2) s.P[j1 * kb + j2] = __fdiv_rn(s.P[j1 * kb + j2], sum);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(ATT_THREADS) attention_fwd_kernel(const AttArgs a) {
  extern __shared__ float sm[];
  const long long node = blockIdx.x;
  const int h = blockIdx.y;
  const int b = (int)a.batch[node];
  const int node0 = a.graph_ptr[b];
  const int n = a.graph_ptr[b + 1] - node0;
  const int kb = a.masked ? (n < a.kslots ? n : a.kslots) : a.kslots;
  const long long r0 = a.row_ptr[b] + (node - node0);
  const int dk = a.dk;
  AttSmem s;
  s.Q = sm; s.K = s.Q + kb * dk; s.V = s.K + kb * dk; s.P = s.V + kb * dk;
  att_load(a, r0, n, kb, h, s);
  __syncthreads();
  att_probs(a, kb, s);
  for (int idx = threadIdx.x; idx < kb * dk; idx += blockDim.x) {
    const int j1 = idx / dk, c = idx - j1 * dk;
    float acc = 0.f;
    for (int j2 = 0; j2 < kb; ++j2) {
      float p = s.P[j1 * kb + j2];
      if (a.drop_p > 0.f) p *= att_keep_scale(a.seed, node, h, j1, j2, a.drop_p);
      acc = fmaf(p, s.V[j2 * dk + c], acc);
    }
    a.o[(r0 + (long long)j1 * n) * a.ld + h * dk + c] = acc;
  }
}

__global__ void __launch_bounds__(ATT_THREADS) attention_bwd_kernel(const AttArgs a) {
  extern __shared__ float sm[];
  const long long node = blockIdx.x;
  const int h = blockIdx.y;
  const int b = (int)a.batch[node];
  const int node0 = a.graph_ptr[b];
  const int n = a.graph_ptr[b + 1] - node0;
  const int kb = a.masked ? (n < a.kslots ? n : a.kslots) : a.kslots;
  const long long r0 = a.row_ptr[b] + (node - node0);
  const int dk = a.dk;
  AttSmem s;
  s.Q = sm; s.K = s.Q + kb * dk; s.V = s.K + kb * dk; s.P = s.V + kb * dk;
  float* dO = s.P + kb * kb;   // [kb][dk]
  float* dS = dO + kb * dk;    // [kb][kb]
  att_load(a, r0, n, kb, h, s);
  for (int idx = threadIdx.x; idx < kb * dk; idx += blockDim.x) {
    const int j = idx / dk, c = idx - j * dk;
    dO[idx] = __ldg(a.go + (r0 + (long long)j * n) * a.ld + h * dk + c);
  }
  __syncthreads();
  att_probs(a, kb, s);
  // dV[j2][c] = sum_j1 Pdrop[j1][j2] dO[j1][c]
  for (int idx = threadIdx.x; idx < kb * dk; idx += blockDim.x) {
    const int j2 = idx / dk, c = idx - j2 * dk;
    float acc = 0.f;
    for (int j1 = 0; j1 < kb; ++j1) {
      float p = s.P[j1 * kb + j2];
      if (a.drop_p > 0.f) p *= att_keep_scale(a.seed, node, h, j1, j2, a.drop_p);
      acc = fmaf(p, dO[j1 * dk + c], acc);
    }
    a.gv[(r0 + (long long)j2 * n) * a.ld + h * dk + c] = acc;
  }
  // dP[j1][j2] = keep * sum_c dO[j1][c] V[j2][c]
  for (int idx = threadIdx.x; idx < kb * kb; idx += blockDim.x) {
    const int j1 = idx / kb, j2 = idx - j1 * kb;
    float acc = 0.f;
    for (int c = 0; c < dk; ++c) acc = fmaf(dO[j1 * dk + c], s.V[j2 * dk + c], acc);
    if (a.drop_p > 0.f) acc *= att_keep_scale(a.seed, node, h, j1, j2, a.drop_p);
    dS[idx] = acc;
  }
  __syncthreads();
  // dS = P * (dP - sum_j2 dP P)
  for (int j1 = threadIdx.x; j1 < kb; j1 += blockDim.x) {
    float dot = 0.f;
    for (int j2 = 0; j2 < kb; ++j2) dot = fmaf(dS[j1 * kb + j2], s.P[j1 * kb + j2], dot);
    for (int j2 = 0; j2 < kb; ++j2) dS[j1 * kb + j2] = s.P[j1 * kb + j2] * (dS[j1 * kb + j2] - dot);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < kb * dk; idx += blockDim.x) {
    const int j = idx / dk, c = idx - j * dk;
    float aq = 0.f, ak = 0.f;
    for (int t = 0; t < kb; ++t) {
      aq = fmaf(dS[j * kb + t], s.K[t * dk + c], aq);   // dQs[j] = sum_t dS[j][t] K[t]
      ak = fmaf(dS[t * kb + j], s.Q[t * dk + c], ak);   // dK[j]  = sum_t dS[t][j] Qs[t]
    }
    const long long off = (r0 + (long long)j * n) * a.ld + h * dk + c;
    a.gq[off] = __fdiv_rn(aq, a.inv_temp_div);
    a.gk[off] = ak;
  }
}

static size_t att_smem(int kmax, int dk, bool bwd) {
  return (size_t)(bwd ? (4 * kmax * dk + 2 * kmax * kmax) : (3 * kmax * dk + kmax * kmax)) * sizeof(float);
}

extern "C" int sb_attention_fwd(const float* q, const float* k, const float* v, int64_t ld, const int64_t* batch,
                                const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N, int32_t kslots,
                                int32_t masked, int32_t kmax, int32_t n_head, int32_t dk, float temperature,
                                float drop_p, int64_t seed, float* o, void* stream) {
  SB_CHECK_ARG(n_head >= 1 && dk >= 1 && kmax >= 1 && ld >= (int64_t)n_head * dk, "sb_attention_fwd: bad sizes");
  if (N == 0) return SB_OK;
  AttArgs a;
  a.q = q; a.k = k; a.v = v; a.ld = ld; a.batch = batch; a.graph_ptr = graph_ptr; a.row_ptr = row_ptr; a.N = N;
  a.kslots = kslots; a.masked = masked; a.n_head = n_head; a.dk = dk; a.inv_temp_div = temperature;
  a.drop_p = drop_p; a.seed = (unsigned long long)seed; a.o = o; a.go = nullptr; a.gq = a.gk = a.gv = nullptr;
  {
    const int rc = sb_attention_mma_fwd_launch(a, kmax, (cudaStream_t)stream);
    if (rc != SB_ERR_UNSUPPORTED) return rc;
  }
  {
    const int rc = sb_attention_fast_fwd_launch(a, kmax, (cudaStream_t)stream);
    if (rc != SB_ERR_UNSUPPORTED) return rc;
  }
  const size_t smem = att_smem(kmax, dk, false);
  SB_CHECK_ARG(smem <= 200 * 1024, "sb_attention_fwd: sequence too long for shared memory (k=%d, dk=%d)", kmax, dk);
  SB_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)N, (unsigned)n_head);
  attention_fwd_kernel<<<grid, ATT_THREADS, smem, (cudaStream_t)stream>>>(a);
  SB_CHECK_LAUNCH("sb_attention_fwd");
  return SB_OK;
}

extern "C" int sb_attention_bwd(const float* q, const float* k, const float* v, const float* go, int64_t ld,
                                const int64_t* batch, const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N,
                                int32_t kslots, int32_t masked, int32_t kmax, int32_t n_head, int32_t dk,
                                float temperature, float drop_p, int64_t seed, float* gq, float* gk, float* gv,
                                void* stream) {
  SB_CHECK_ARG(n_head >= 1 && dk >= 1 && kmax >= 1 && ld >= (int64_t)n_head * dk, "sb_attention_bwd: bad sizes");
  if (N == 0) return SB_OK;
  AttArgs a;
  a.q = q; a.k = k; a.v = v; a.ld = ld; a.batch = batch; a.graph_ptr = graph_ptr; a.row_ptr = row_ptr; a.N = N;
  a.kslots = kslots; a.masked = masked; a.n_head = n_head; a.dk = dk; a.inv_temp_div = temperature;
  a.drop_p = drop_p; a.seed = (unsigned long long)seed; a.o = nullptr; a.go = go; a.gq = gq; a.gk = gk; a.gv = gv;
  {
    const int rc = sb_attention_mma_bwd_launch(a, kmax, (cudaStream_t)stream);
    if (rc != SB_ERR_UNSUPPORTED) return rc;
  }
  {
    const int rc = sb_attention_fast_bwd_launch(a, kmax, (cudaStream_t)stream);
    if (rc != SB_ERR_UNSUPPORTED) return rc;
  }
  const size_t smem = att_smem(kmax, dk, true);
  SB_CHECK_ARG(smem <= 200 * 1024, "sb_attention_bwd: sequence too long for shared memory (k=%d, dk=%d)", kmax, dk);
  SB_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)N, (unsigned)n_head);
  attention_bwd_kernel<<<grid, ATT_THREADS, smem, (cudaStream_t)stream>>>(a);
  SB_CHECK_LAUNCH("sb_attention_bwd");
  return SB_OK;
}

// --------------------------------------------------------------------------------------------------------- LayerNorm
// y = LN(a + b) * w + beta  per row (biased variance, eps given; MaskedLN uses 1e-6), one warp per row.
// stat[r] = (mean, rstd) kept for the backward; sum_out = a + b is written when requested (LN input, for the backward).
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ w,
                                                            const float* __restrict__ beta, long long ld, long long R,
                                                            int C, float eps, float* __restrict__ y,
                                                            float* __restrict__ xsum, float* __restrict__ stat) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  float v[8];  // C <= 256
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    float t = 0.f;
    if (c < C) {
      t = __ldg(a + row * ld + c);
      if (b) t += __ldg(b + row * ld + c);
    }
    v[i] = t;
    s += t;
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c < C) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  }
  const float var = warp_sum(q) / (float)C;
  const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c < (int)ld) {
      float o = 0.f;
      if (c < C) o = (v[i] - mean) * rstd * __ldg(w + c) + __ldg(beta + c);
      y[row * ld + c] = o;
      if (xsum) xsum[row * ld + c] = (c < C) ? v[i] : 0.f;
    }
  }
  if (lane == 0 && stat) { stat[row * 2] = mean; stat[row * 2 + 1] = rstd; }
}
// Vectorised variant (C % 4 == 0, C <= 128, 16-byte aligned rows): a warp takes four rows at a time, a lane owns one
// float4 column group, all eight 128-bit loads are issued before the first reduction (the scalar kernel kept 32 B per
// thread in flight and ran at 0.54 of the HBM peak).
#define LN_ROWS 4
__global__ void __launch_bounds__(256) layernorm_fwd_vec_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ beta, long long ld, long long R,
                                                                int C, float eps, float* __restrict__ y,
                                                                float* __restrict__ xsum, float* __restrict__ stat) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int col = lane * 4;
  const bool in_ld = col < (int)ld, in_c = col < C;
  float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), be4 = w4;
  if (in_c) { w4 = ldg4(w + col); be4 = ldg4(beta + col); }
  const float invC = 1.0f / (float)C;
  for (long long r0 = warp0 * LN_ROWS; r0 < R; r0 += nwarps * LN_ROWS) {
    float4 v[LN_ROWS];
#pragma unroll
    for (int k = 0; k < LN_ROWS; ++k) {
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in_c && r0 + k < R) {
        v[k] = ldg4(a + (r0 + k) * ld + col);
        if (b) {
          const float4 t = ldg4(b + (r0 + k) * ld + col);
          v[k].x += t.x; v[k].y += t.y; v[k].z += t.z; v[k].w += t.w;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < LN_ROWS; ++k) {
      if (r0 + k >= R) break;   // warp-uniform
      const float mean = warp_sum((v[k].x + v[k].y) + (v[k].z + v[k].w)) * invC;
      float q = 0.f;
      if (in_c) {
        const float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
        q = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
      }
      const float var = warp_sum(q) * invC;
      const float rstd = 1.0f / sqrtf(var + eps);
      if (in_ld) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_c) {
          o.x = (v[k].x - mean) * rstd * w4.x + be4.x;
          o.y = (v[k].y - mean) * rstd * w4.y + be4.y;
          o.z = (v[k].z - mean) * rstd * w4.z + be4.z;
          o.w = (v[k].w - mean) * rstd * w4.w + be4.w;
        }
        *reinterpret_cast<float4*>(y + (r0 + k) * ld + col) = o;
        if (xsum) *reinterpret_cast<float4*>(xsum + (r0 + k) * ld + col) = v[k];
      }
      if (lane == 0 && stat) { stat[(r0 + k) * 2] = mean; stat[(r0 + k) * 2 + 1] = rstd; }
    }
  }
}
static bool ln_vec_ok(const void* p0, const void* p1, const void* p2, const void* p3, long long ld, int C) {
  auto al = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; };
  return C % 4 == 0 && C <= 128 && ld % 4 == 0 && ld <= 128 && al(p0) && al(p1) && al(p2) && al(p3);
}
extern "C" int sb_layernorm_fwd(const float* a, const float* b, const float* w, const float* beta, int64_t ld,
                                int64_t R, int32_t C, float eps, float* y, float* xsum, float* stat, void* stream) {
  SB_CHECK_ARG(C >= 1 && C <= 256 && ld >= C && ld <= 256, "sb_layernorm_fwd: feature dim must be <= 256");
  if (R == 0) return SB_OK;
  if (ln_vec_ok(a, b, y, xsum, ld, C) && ((uintptr_t)w & 15u) == 0 && ((uintptr_t)beta & 15u) == 0) {
    long long blocks = sb_ceil_div(R, 8 * LN_ROWS);
    const long long cap = (long long)sb_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    layernorm_fwd_vec_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, w, beta, ld, R, C, eps, y, xsum,
                                                                                stat);
    SB_CHECK_LAUNCH("sb_layernorm_fwd(vec)");
    return SB_OK;
  }
  layernorm_fwd_kernel<<<(unsigned)sb_ceil_div(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(a, b, w, beta, ld, R, C,
                                                                                           eps, y, xsum, stat);
  SB_CHECK_LAUNCH("sb_layernorm_fwd");
  return SB_OK;
}

// dx = rstd * (g*w - mean_c(g*w) - xhat * mean_c(g*w*xhat));  dw += sum_r g*xhat;  dbeta += sum_r g   (fp64 partials)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                            const float* __restrict__ stat,
                                                            const float* __restrict__ w, long long ld, long long R,
                                                            int C, float* __restrict__ dx, double* __restrict__ dwb) {
  __shared__ double red[2][256];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  double dw[8], db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dw[i] = db[i] = 0.0;
  for (long long row = (long long)blockIdx.x * 8 + wrp; row < R; row += (long long)gridDim.x * 8) {
    const float mean = stat[row * 2], rstd = stat[row * 2 + 1];
    float gw[8], xh[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      gw[i] = xh[i] = 0.f;
      if (c < C) {
        const float gv = __ldg(g + row * ld + c);
        xh[i] = (__ldg(x + row * ld + c) - mean) * rstd;
        gw[i] = gv * __ldg(w + c);
        s1 += gw[i];
        s2 = fmaf(gw[i], xh[i], s2);
        dw[i] += (double)gv * (double)xh[i];
        db[i] += (double)gv;
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < (int)ld) dx[row * ld + c] = (c < C) ? rstd * (gw[i] - s1 - xh[i] * s2) : 0.f;
    }
  }
  // reduce the per-warp column partials over the 8 warps in shared memory, then one fp64 atomic per column per CTA
  for (int c = threadIdx.x; c < 256; c += blockDim.x) { red[0][c] = 0.0; red[1][c] = 0.0; }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + 32 * i;
    if (c < C) { atomicAdd(&red[0][c], dw[i]); atomicAdd(&red[1][c], db[i]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(dwb + c, red[0][c]);
    atomicAdd(dwb + C + c, red[1][c]);
  }
}
__global__ void __launch_bounds__(256) layernorm_bwd_vec_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                                const float* __restrict__ stat,
                                                                const float* __restrict__ w, long long ld, long long R,
                                                                int C, float* __restrict__ dx, double* __restrict__ dwb) {
  __shared__ double red[2][128];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int col = lane * 4;
  const bool in_ld = col < (int)ld, in_c = col < C;
  float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (in_c) w4 = ldg4(w + col);
  const float invC = 1.0f / (float)C;
  double dw[4] = {0.0, 0.0, 0.0, 0.0}, db[4] = {0.0, 0.0, 0.0, 0.0};
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long r0 = ((long long)blockIdx.x * 8 + wrp) * LN_ROWS; r0 < R; r0 += nwarps * LN_ROWS) {
    float4 gv[LN_ROWS], xv[LN_ROWS];
    float2 st[LN_ROWS];
#pragma unroll
    for (int k = 0; k < LN_ROWS; ++k) {
      gv[k] = xv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      st[k] = make_float2(0.f, 0.f);
      if (r0 + k < R) {
        st[k] = __ldg(reinterpret_cast<const float2*>(stat + (r0 + k) * 2));
        if (in_c) {
          gv[k] = ldg4(g + (r0 + k) * ld + col);
          xv[k] = ldg4(x + (r0 + k) * ld + col);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < LN_ROWS; ++k) {
      if (r0 + k >= R) break;   // warp-uniform
      const float mean = st[k].x, rstd = st[k].y;
      const float gi[4] = {gv[k].x, gv[k].y, gv[k].z, gv[k].w};
      const float xi[4] = {xv[k].x, xv[k].y, xv[k].z, xv[k].w};
      const float wi[4] = {w4.x, w4.y, w4.z, w4.w};
      float gw[4], xh[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xh[j] = in_c ? (xi[j] - mean) * rstd : 0.f;
        gw[j] = gi[j] * wi[j];
        s1 += gw[j];
        s2 = fmaf(gw[j], xh[j], s2);
        dw[j] += (double)gi[j] * (double)xh[j];
        db[j] += (double)gi[j];
      }
      s1 = warp_sum(s1) * invC;
      s2 = warp_sum(s2) * invC;
      if (in_ld) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_c) {
          o.x = rstd * (gw[0] - s1 - xh[0] * s2);
          o.y = rstd * (gw[1] - s1 - xh[1] * s2);
          o.z = rstd * (gw[2] - s1 - xh[2] * s2);
          o.w = rstd * (gw[3] - s1 - xh[3] * s2);
        }
        *reinterpret_cast<float4*>(dx + (r0 + k) * ld + col) = o;
      }
    }
  }
  for (int c = threadIdx.x; c < 128; c += blockDim.x) { red[0][c] = 0.0; red[1][c] = 0.0; }
  __syncthreads();
  if (in_c) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&red[0][col + j], dw[j]); atomicAdd(&red[1][col + j], db[j]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(dwb + c, red[0][c]);
    atomicAdd(dwb + C + c, red[1][c]);
  }
}
extern "C" int sb_layernorm_bwd(const float* g, const float* x, const float* stat, const float* w, int64_t ld,
                                int64_t R, int32_t C, float* dx, double* dwb, void* stream) {
  SB_CHECK_ARG(C >= 1 && C <= 256 && ld >= C && ld <= 256, "sb_layernorm_bwd: feature dim must be <= 256");
  if (R == 0) return SB_OK;
  long long blocks = sb_ceil_div(R, 8 * 4);
  const long long cap = (long long)sb_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (ln_vec_ok(g, x, dx, w, ld, C)) {
    layernorm_bwd_vec_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, x, stat, w, ld, R, C, dx, dwb);
    SB_CHECK_LAUNCH("sb_layernorm_bwd(vec)");
    return SB_OK;
  }
  layernorm_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, x, stat, w, ld, R, C, dx, dwb);
  SB_CHECK_LAUNCH("sb_layernorm_bwd");
  return SB_OK;
}

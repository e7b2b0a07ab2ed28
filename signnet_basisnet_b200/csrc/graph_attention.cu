// K13 — edge-modulated sparse attention of the graph-Transformer predictor (GraphPrediction/layers/transformer.py:160-192,
// MultiHeadAttentionLayer.propagate_attention with full_graph=False) on [N, ld] node rows / [E, ld] edge rows, H heads of
// width d (H*d <= ld).  Per edge k: j -> i and head h:
//     a_k = sum_c ((K[j,c] * Q[i,c]) / sqrt(d)) * E[k,c]          s_k = exp(clamp(a_k, -5, 5))
//     out[i, h] = (sum_in s_k V[j, h]) / (sum_in s_k + 1e-6)
// One THREAD per (destination node, head): the head's d <= 32 channels are walked serially, incoming edges in stable CSR
// (= edge id) order - deterministic, no atomics, no inter-thread communication (tests/test_cpu_emulation_attention.py
// runs this source text thread by thread on the CPU).  The raw scores a_k [E, H] and the normalisers z [N, H] are kept
// for the backward, which writes per-edge gradients (dE, and the K / V contributions that a second pass sums per source).
// STATUS: written after the round's GPU budget was spent; compiles for sm_100a, not yet run on a GPU.
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define GA_MAXD 32

__global__ void __launch_bounds__(256) edge_attention_fwd_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                                 const float* __restrict__ Ef, const float* __restrict__ V,
                                                                 const int32_t* __restrict__ in_ptr,
                                                                 const int32_t* __restrict__ in_src,
                                                                 const int32_t* __restrict__ in_eid, long long N, int H,
                                                                 int d, long long ld, float* __restrict__ out,
                                                                 float* __restrict__ araw, float* __restrict__ z) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= N * H) return;
  const long long node = idx / H;
  const int head = (int)(idx - node * H);
  const int c0 = head * d;
  const float rs = sqrtf((float)d);
  const int beg = __ldg(in_ptr + node), end = __ldg(in_ptr + node + 1);
  float acc[GA_MAXD];
#pragma unroll
  for (int c = 0; c < GA_MAXD; ++c) acc[c] = 0.f;
  float zs = 0.f;
  for (int p = beg; p < end; ++p) {
    const long long j = __ldg(in_src + p), k = __ldg(in_eid + p);
    float a = 0.f;
    for (int c = 0; c < d; ++c)
      a = __fadd_rn(a, __fmul_rn(__fdiv_rn(__fmul_rn(__ldg(K + j * ld + c0 + c), __ldg(Q + node * ld + c0 + c)), rs),
                                 __ldg(Ef + k * ld + c0 + c)));
    araw[k * H + head] = a;
    const float s = expf(fminf(fmaxf(a, -5.f), 5.f));
    for (int c = 0; c < d; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(__ldg(V + j * ld + c0 + c), s));
    zs = __fadd_rn(zs, s);
  }
  z[node * H + head] = zs;
  const float den = __fadd_rn(zs, 1e-6f);
  for (int c = 0; c < d; ++c) out[node * ld + c0 + c] = __fdiv_rn(acc[c], den);
  if (head == 0)
    for (long long c = (long long)H * d; c < ld; ++c) out[node * ld + c] = 0.f;
}

// backward, destination part (thread per (node i, head)): dQ[i], and per incoming edge dE[k], dKe[k], dVe[k]
__global__ void __launch_bounds__(256) edge_attention_bwd_dst_kernel(
    const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ Q, const float* __restrict__ K,
    const float* __restrict__ Ef, const float* __restrict__ V, const float* __restrict__ araw, const float* __restrict__ z,
    const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src, const int32_t* __restrict__ in_eid, long long N,
    int H, int d, long long ld, float* __restrict__ dQ, float* __restrict__ dE, float* __restrict__ dKe,
    float* __restrict__ dVe) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= N * H) return;
  const long long node = idx / H;
  const int head = (int)(idx - node * H);
  const int c0 = head * d;
  const float rs = sqrtf((float)d);
  const int beg = __ldg(in_ptr + node), end = __ldg(in_ptr + node + 1);
  const float r = 1.0f / (__ldg(z + node * H + head) + 1e-6f);
  float dwv[GA_MAXD], dq[GA_MAXD];
  float dz = 0.f;
#pragma unroll
  for (int c = 0; c < GA_MAXD; ++c) { dwv[c] = 0.f; dq[c] = 0.f; }
  for (int c = 0; c < d; ++c) {
    const float g = __ldg(dout + node * ld + c0 + c);
    dwv[c] = g * r;
    dz -= g * __ldg(out + node * ld + c0 + c) * r;   // out = wV r  ->  d out / d z = -wV r^2 = -out r
  }
  for (int p = beg; p < end; ++p) {
    const long long j = __ldg(in_src + p), k = __ldg(in_eid + p);
    const float a = __ldg(araw + k * H + head);
    const float s = expf(fminf(fmaxf(a, -5.f), 5.f));
    float ds = dz;
    for (int c = 0; c < d; ++c) {
      ds += dwv[c] * __ldg(V + j * ld + c0 + c);
      dVe[k * ld + c0 + c] = s * dwv[c];
    }
    const float da = (a >= -5.f && a <= 5.f) ? ds * s : 0.f;   // clamp passes the gradient inside [-5, 5]
    for (int c = 0; c < d; ++c) {
      const float kv = __ldg(K + j * ld + c0 + c), qv = __ldg(Q + node * ld + c0 + c), ev = __ldg(Ef + k * ld + c0 + c);
      dq[c] += da * kv * ev / rs;
      dKe[k * ld + c0 + c] = da * qv * ev / rs;
      dE[k * ld + c0 + c] = da * kv * qv / rs;
    }
    if (head == 0)
      for (long long c = (long long)H * d; c < ld; ++c) { dE[k * ld + c] = 0.f; dKe[k * ld + c] = 0.f; dVe[k * ld + c] = 0.f; }
  }
  for (int c = 0; c < d; ++c) dQ[node * ld + c0 + c] = dq[c];
  if (head == 0)
    for (long long c = (long long)H * d; c < ld; ++c) dQ[node * ld + c] = 0.f;
}
// backward, source part (thread per (node j, column)): dK[j] = sum over outgoing edges of dKe, dV[j] likewise (CSC order)
__global__ void __launch_bounds__(256) edge_attention_bwd_src_kernel(const float* __restrict__ dKe,
                                                                     const float* __restrict__ dVe,
                                                                     const int32_t* __restrict__ out_ptr,
                                                                     const int32_t* __restrict__ out_eid, long long N,
                                                                     long long ld, float* __restrict__ dK,
                                                                     float* __restrict__ dV) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= N * ld) return;
  const long long node = idx / ld, c = idx - node * ld;
  const int beg = __ldg(out_ptr + node), end = __ldg(out_ptr + node + 1);
  float ak = 0.f, av = 0.f;
  for (int p = beg; p < end; ++p) {
    const long long k = __ldg(out_eid + p);
    ak += __ldg(dKe + k * ld + c);
    av += __ldg(dVe + k * ld + c);
  }
  dK[idx] = ak;
  dV[idx] = av;
}

extern "C" int sb_edge_attention_fwd(const float* Q, const float* K, const float* Ef, const float* V, const int32_t* in_ptr,
                                     const int32_t* in_src, const int32_t* in_eid, int64_t N, int32_t H, int32_t d,
                                     int64_t ld, float* out, float* araw, float* z, void* stream) {
  SB_CHECK_ARG(H >= 1 && d >= 1 && d <= GA_MAXD && (int64_t)H * d <= ld, "sb_edge_attention_fwd: bad sizes H=%d d=%d", H, d);
  if (N == 0) return SB_OK;
  edge_attention_fwd_kernel<<<(unsigned)sb_ceil_div(N * H, 256), 256, 0, (cudaStream_t)stream>>>(Q, K, Ef, V, in_ptr, in_src,
                                                                                               in_eid, N, H, d, ld, out, araw, z);
  SB_CHECK_LAUNCH("sb_edge_attention_fwd");
  return SB_OK;
}

extern "C" int sb_edge_attention_bwd(const float* dout, const float* out, const float* Q, const float* K, const float* Ef,
                                     const float* V, const float* araw, const float* z, const int32_t* in_ptr,
                                     const int32_t* in_src, const int32_t* in_eid, const int32_t* out_ptr,
                                     const int32_t* out_eid, int64_t N, int32_t H, int32_t d, int64_t ld, float* dQ,
                                     float* dK, float* dE, float* dV, float* dKe, float* dVe, void* stream) {
  SB_CHECK_ARG(H >= 1 && d >= 1 && d <= GA_MAXD && (int64_t)H * d <= ld, "sb_edge_attention_bwd: bad sizes H=%d d=%d", H, d);
  if (N == 0) return SB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  edge_attention_bwd_dst_kernel<<<(unsigned)sb_ceil_div(N * H, 256), 256, 0, st>>>(dout, out, Q, K, Ef, V, araw, z, in_ptr, in_src,
                                                                                 in_eid, N, H, d, ld, dQ, dE, dKe, dVe);
  SB_CHECK_LAUNCH("sb_edge_attention_bwd(dst)");
  edge_attention_bwd_src_kernel<<<(unsigned)sb_ceil_div(N * ld, 256), 256, 0, st>>>(dKe, dVe, out_ptr, out_eid, N, ld, dK, dV);
  SB_CHECK_LAUNCH("sb_edge_attention_bwd(src)");
  return SB_OK;
}

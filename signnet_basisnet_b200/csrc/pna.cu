// K12 — the multi-aggregator neighbourhood reduction of the PNA predictor (GraphPrediction/layers/pna_layer.py:37-68,
// pna_utils.py:12-31,73-84) on [N, ld] node rows and [E, ld] edge rows.  The tower's pre-transformation
// Linear(cat[h_src, h_dst, e]) is linear in its three inputs, so the message of edge k: j -> i is
//     m_k = (U[j] + V[i]) + Q[k]            U = h W_src^T, V = h W_dst^T (node-level Linears), Q = e W_e^T + b (edge-level)
// and is never materialised.  Per destination node and channel, over its D incoming messages:
//     mean, max, min, std = sqrt(relu(E[m^2] - E[m]^2) + 1e-5)          (aggregators "mean max min std")
//     x 1, x log(D+1)/avg, x avg/log(D+1)                                (scalers "identity amplification attenuation")
// written tower-major next to the node's own features, which is the input of the post-transformation Linear:
//     Z[i, t*13*tin + 0 .. tin)                      = h[i, t*tin .. (t+1)*tin)
//     Z[i, t*13*tin + tin + (s*4 + a)*tin + j]       = scaler s of aggregator a of channel t*tin + j
// Nodes without incoming edges get zeros (DGL does not call the reduce function for them).
// One warp per node, lanes over channels, incoming edges in stable CSR (= edge id) order; deterministic, no atomics, no
// inter-thread communication (tests/test_cpu_emulation_pna.py runs this source text thread by thread on the CPU).
// STATUS: written after the round's GPU budget was spent; compiles for sm_100a, not yet run on a GPU.
#include "common.cuh"
#include "../../include/signnet_b200.h"

__global__ void __launch_bounds__(256) pna_agg_fwd_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                                          const float* __restrict__ Q, const float* __restrict__ h,
                                                          const int32_t* __restrict__ in_ptr,
                                                          const int32_t* __restrict__ in_src,
                                                          const int32_t* __restrict__ in_eid, long long N, int C, int tin,
                                                          long long ld, long long ldh, long long ldz, float avg_log,
                                                          float* __restrict__ Z) {
  const long long node = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= N) return;
  const int beg = __ldg(in_ptr + node), end = __ldg(in_ptr + node + 1);
  const int D = end - beg;
  const float logd = logf((float)(D > 0 ? D : 1) + 1.0f);
  const float scale[3] = {1.0f, logd / avg_log, avg_log / logd};
  for (int c = lane; c < C; c += 32) {
    const float v = __ldg(V + node * ld + c);
    float s1 = 0.f, s2 = 0.f, mx = -INFINITY, mn = INFINITY;
    for (int p = beg; p < end; ++p) {
      const float m = __fadd_rn(__fadd_rn(__ldg(U + (long long)__ldg(in_src + p) * ld + c), v),
                                __ldg(Q + (long long)__ldg(in_eid + p) * ld + c));
      s1 = __fadd_rn(s1, m);
      s2 = __fadd_rn(s2, __fmul_rn(m, m));
      mx = fmaxf(mx, m);
      mn = fminf(mn, m);
    }
    const int t = c / tin, j = c - t * tin;
    float* z = Z + node * ldz + (long long)t * 13 * tin;
    z[j] = __ldg(h + node * ldh + c);
    float agg[4] = {0.f, 0.f, 0.f, 0.f};
    if (D > 0) {
      const float mean = s1 / (float)D;
      const float var = fmaxf(__fadd_rn(s2 / (float)D, -__fmul_rn(mean, mean)), 0.f);
      agg[0] = mean; agg[1] = mx; agg[2] = mn; agg[3] = sqrtf(var + 1e-5f);
    }
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int a = 0; a < 4; ++a) z[tin + (s * 4 + a) * tin + j] = (D > 0) ? agg[a] * scale[s] : 0.f;
  }
  // zero the padding columns of the row once (lane 0 .. covers [13 C, ldz))
  for (long long c = 13ll * C + lane; c < ldz; c += 32) Z[node * ldz + c] = 0.f;
}

// backward, destination-node part: recompute the statistics, route the gradient of the 12 aggregate columns to every
// incoming message (max / min go to the FIRST edge attaining them, like torch.max / torch.min), write it per edge
// (dQ = dm) and sum it per destination (dV); dh takes the pass-through columns.
__global__ void __launch_bounds__(256) pna_agg_bwd_dst_kernel(const float* __restrict__ dZ, const float* __restrict__ U,
                                                              const float* __restrict__ V, const float* __restrict__ Q,
                                                              const int32_t* __restrict__ in_ptr,
                                                              const int32_t* __restrict__ in_src,
                                                              const int32_t* __restrict__ in_eid, long long N, int C,
                                                              int tin, long long ld, long long ldh, long long ldz,
                                                              float avg_log, float* __restrict__ dV,
                                                              float* __restrict__ dQ, float* __restrict__ dh) {
  const long long node = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= N) return;
  const int beg = __ldg(in_ptr + node), end = __ldg(in_ptr + node + 1);
  const int D = end - beg;
  const float logd = logf((float)(D > 0 ? D : 1) + 1.0f);
  const float scale[3] = {1.0f, logd / avg_log, avg_log / logd};
  for (int c = lane; c < C; c += 32) {
    const int t = c / tin, j = c - t * tin;
    const float* gz = dZ + node * ldz + (long long)t * 13 * tin;
    dh[node * ldh + c] = gz[j];
    const float v = __ldg(V + node * ld + c);
    float s1 = 0.f, s2 = 0.f, mx = -INFINITY, mn = INFINITY;
    int pmx = beg, pmn = beg;
    for (int p = beg; p < end; ++p) {
      const float m = __fadd_rn(__fadd_rn(__ldg(U + (long long)__ldg(in_src + p) * ld + c), v),
                                __ldg(Q + (long long)__ldg(in_eid + p) * ld + c));
      s1 = __fadd_rn(s1, m);
      s2 = __fadd_rn(s2, __fmul_rn(m, m));
      if (m > mx) { mx = m; pmx = p; }
      if (m < mn) { mn = m; pmn = p; }
    }
    float acc = 0.f;
    if (D > 0) {
      float g[4] = {0.f, 0.f, 0.f, 0.f};   // gradient w.r.t. mean, max, min, std (scalers folded in)
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int a = 0; a < 4; ++a) g[a] += scale[s] * gz[tin + (s * 4 + a) * tin + j];
      const float invD = 1.0f / (float)D;
      const float mean = s1 * invD;
      const float var = __fadd_rn(s2 * invD, -__fmul_rn(mean, mean));
      const float stdv = sqrtf(fmaxf(var, 0.f) + 1e-5f);
      const float gvar = (var > 0.f) ? g[3] / (2.0f * stdv) : 0.f;   // d std / d var through relu
      for (int p = beg; p < end; ++p) {
        const long long k = __ldg(in_eid + p);
        const float m = __fadd_rn(__fadd_rn(__ldg(U + (long long)__ldg(in_src + p) * ld + c), v), __ldg(Q + k * ld + c));
        float dm = g[0] * invD + gvar * 2.0f * invD * (m - mean);
        if (p == pmx) dm += g[1];
        if (p == pmn) dm += g[2];
        dQ[k * ld + c] = dm;
        acc += dm;
      }
    }
    dV[node * ld + c] = acc;
  }
  for (long long c = C + lane; c < ld; c += 32) dV[node * ld + c] = 0.f;
  for (long long c = C + lane; c < ldh; c += 32) dh[node * ldh + c] = 0.f;
  for (int p = beg; p < end; ++p) {   // padding columns of this node's incoming edge rows
    const long long k = __ldg(in_eid + p);
    for (long long c = C + lane; c < ld; c += 32) dQ[k * ld + c] = 0.f;
  }
}
// backward, source-node part: dU_j = sum over the outgoing edges of j of dm (fixed CSC order)
__global__ void __launch_bounds__(256) pna_agg_bwd_src_kernel(const float* __restrict__ dQ,
                                                              const int32_t* __restrict__ out_ptr,
                                                              const int32_t* __restrict__ out_eid, long long N,
                                                              long long ld, float* __restrict__ dU) {
  const long long node = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= N) return;
  const int beg = __ldg(out_ptr + node), end = __ldg(out_ptr + node + 1);
  for (long long c = lane; c < ld; c += 32) {
    float acc = 0.f;
    for (int p = beg; p < end; ++p) acc += __ldg(dQ + (long long)__ldg(out_eid + p) * ld + c);
    dU[node * ld + c] = acc;
  }
}

extern "C" int sb_pna_agg_fwd(const float* U, const float* V, const float* Q, const float* h, const int32_t* in_ptr,
                              const int32_t* in_src, const int32_t* in_eid, int64_t N, int32_t C, int32_t tin, int64_t ld,
                              int64_t ldh, int64_t ldz, float avg_log, float* Z, void* stream) {
  SB_CHECK_ARG(C >= 1 && tin >= 1 && C % tin == 0 && ld >= C && ldh >= C && ldz >= 13ll * C && avg_log > 0.f,
               "sb_pna_agg_fwd: bad sizes C=%d tin=%d", C, tin);
  if (N == 0) return SB_OK;
  pna_agg_fwd_kernel<<<(unsigned)sb_ceil_div(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(U, V, Q, h, in_ptr, in_src, in_eid,
                                                                                         N, C, tin, ld, ldh, ldz, avg_log, Z);
  SB_CHECK_LAUNCH("sb_pna_agg_fwd");
  return SB_OK;
}

extern "C" int sb_pna_agg_bwd(const float* dZ, const float* U, const float* V, const float* Q, const int32_t* in_ptr,
                              const int32_t* in_src, const int32_t* in_eid, const int32_t* out_ptr,
                              const int32_t* out_eid, int64_t N, int32_t C, int32_t tin, int64_t ld, int64_t ldh,
                              int64_t ldz, float avg_log, float* dU, float* dV, float* dQ, float* dh, void* stream) {
  SB_CHECK_ARG(C >= 1 && tin >= 1 && C % tin == 0 && ld >= C && ldh >= C && ldz >= 13ll * C && avg_log > 0.f,
               "sb_pna_agg_bwd: bad sizes C=%d tin=%d", C, tin);
  if (N == 0) return SB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  pna_agg_bwd_dst_kernel<<<(unsigned)sb_ceil_div(N * 32, 256), 256, 0, st>>>(dZ, U, V, Q, in_ptr, in_src, in_eid, N, C, tin, ld,
                                                                           ldh, ldz, avg_log, dV, dQ, dh);
  SB_CHECK_LAUNCH("sb_pna_agg_bwd(dst)");
  pna_agg_bwd_src_kernel<<<(unsigned)sb_ceil_div(N * 32, 256), 256, 0, st>>>(dQ, out_ptr, out_eid, N, ld, dU);
  SB_CHECK_LAUNCH("sb_pna_agg_bwd(src)");
  return SB_OK;
}

// ---- element-wise companions of the PNA layer: graph normalisation h * snorm_n (pna_layer.py:73-74) and the
// LeakyReLU(0.01) of the mixing network (pna_utils.py FCLayer, activation='LeakyReLU').  Thread per element.
__global__ void __launch_bounds__(256) row_scale_kernel(const float* __restrict__ x, const float* __restrict__ s, long long M,
                                                        long long ld, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= M * ld) return;
  out[i] = __fmul_rn(x[i], __ldg(s + i / ld));
}
__global__ void __launch_bounds__(256) leaky_relu_kernel(const float* __restrict__ g, const float* __restrict__ x, long long n,
                                                         float slope, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = g ? g[i] : x[i];          // forward: g == NULL -> act(x); backward: g * act'(x)
  out[i] = (x[i] > 0.f) ? v : __fmul_rn(v, slope);
}

/* out[r, :] = x[r, :] * s[r]  (its own backward with the gradient as x) */
extern "C" int sb_row_scale(const float* x, const float* s, int64_t M, int64_t ld, float* out, void* stream) {
  SB_CHECK_ARG(M >= 0 && ld >= 1, "sb_row_scale: bad sizes");
  if (M == 0) return SB_OK;
  row_scale_kernel<<<(unsigned)sb_ceil_div(M * ld, 256), 256, 0, (cudaStream_t)stream>>>(x, s, M, ld, out);
  SB_CHECK_LAUNCH("sb_row_scale");
  return SB_OK;
}
/* g == NULL: out = leaky_relu(x, slope);  else: out = g * (x > 0 ? 1 : slope) */
extern "C" int sb_leaky_relu(const float* g, const float* x, int64_t n, float slope, float* out, void* stream) {
  SB_CHECK_ARG(n >= 0, "sb_leaky_relu: bad size");
  if (n == 0) return SB_OK;
  leaky_relu_kernel<<<(unsigned)sb_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(g, x, n, slope, out);
  SB_CHECK_LAUNCH("sb_leaky_relu");
  return SB_OK;
}

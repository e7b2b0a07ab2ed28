// K7 fast path — self-attention of a node over its k_b <= 40 eigenvector-slot tokens with d_k <= 32, one WARP per
// (node, head): a lane owns one query row (scores, probabilities and the output row stay in registers), K/V rows are
// broadcast out of shared memory; the backward follows the flash-attention recipe (recompute P from Q, K and the saved
// row statistics instead of storing k x k matrices): phase A, lane = query (dQ, row sums D), phase B, lane = key
// (dK, dV).  Semantics identical to the generic kernels in transformer.cu
// (Alchemy/sign_net/model_utils/transformer_module.py:44-58,76-102).
#include "attention.cuh"
#include "../../include/signnet_b200.h"

#define AF_KMAX 40
#define AF_DK 32
#define AF_WARPS 4        // forward: warps (heads) per CTA
#define AF_WARPS_B 2      // backward
#define AF_LDS 36         // padded row stride (floats): rows stay 16-byte aligned for float4 broadcast reads

struct AfCtx {
  long long r0;
  int n, kb, h, dk, dk4;
  bool active;
};

template <int WARPS>
__device__ __forceinline__ AfCtx af_ctx(const AttArgs& a) {
  AfCtx c;
  const long long node = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  c.h = blockIdx.y * WARPS + warp;
  c.active = c.h < a.n_head;
  const int b = (int)a.batch[node];
  const int node0 = a.graph_ptr[b];
  c.n = a.graph_ptr[b + 1] - node0;
  c.kb = a.masked ? (c.n < a.kslots ? c.n : a.kslots) : a.kslots;
  c.r0 = a.row_ptr[b] + (node - node0);
  c.dk = a.dk;
  c.dk4 = (a.dk + 3) >> 2;
  return c;
}
// cooperative (one warp) load of the head slice of every token row into zero-padded shared rows.  All loads of the
// warp are issued before the first dependent use (fixed trip count, registers), otherwise every iteration would
// expose a full DRAM round trip.
__device__ __forceinline__ void af_load(const float* __restrict__ src, long long ld, const AfCtx& c, float* dst,
                                        float div) {
  const int lane = threadIdx.x & 31;
  if ((c.dk & 3) == 0 && (ld & 3) == 0) {
    constexpr int ITERS = (AF_KMAX * (AF_DK / 4) + 31) / 32;   // 10
    const int w4 = c.dk4, total = c.kb * w4;
    float4 v[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int idx = lane + 32 * i;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < total) {
        const int j = idx / w4, c4 = idx - j * w4;
        v[i] = ldg4(src + (c.r0 + (long long)j * c.n) * ld + c.h * c.dk + c4 * 4);
      }
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int idx = lane + 32 * i;
      if (idx < total) {
        const int j = idx / w4, c4 = idx - j * w4;
        float4 t = v[i];
        if (div != 0.f) {
          t.x = __fdiv_rn(t.x, div); t.y = __fdiv_rn(t.y, div); t.z = __fdiv_rn(t.z, div); t.w = __fdiv_rn(t.w, div);
        }
        *reinterpret_cast<float4*>(dst + j * AF_LDS + c4 * 4) = t;
      }
    }
    return;
  }
  const int w = c.dk4 * 4;
  for (int idx = lane; idx < c.kb * w; idx += 32) {
    const int j = idx / w, cc = idx - j * w;
    float v = 0.f;
    if (cc < c.dk) {
      v = __ldg(src + (c.r0 + (long long)j * c.n) * ld + c.h * c.dk + cc);
      if (div != 0.f) v = __fdiv_rn(v, div);
    }
    dst[j * AF_LDS + cc] = v;
  }
}
// Asynchronous variant (cp.async, 16 B per op, no registers): every operand tile of the warp is in flight at once.
__device__ __forceinline__ bool af_vec_ok(const AttArgs& a) {
  return (a.dk & 3) == 0 && (a.ld & 3) == 0;
}
__device__ __forceinline__ void af_load_async(const float* __restrict__ src, long long ld, const AfCtx& c, float* dst) {
  const int lane = threadIdx.x & 31;
  const int w4 = c.dk4, total = c.kb * w4;
  for (int idx = lane; idx < total; idx += 32) {
    const int j = idx / w4, c4 = idx - j * w4;
    const float* g = src + (c.r0 + (long long)j * c.n) * ld + c.h * c.dk + c4 * 4;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + j * AF_LDS + c4 * 4)), "l"(g) : "memory");
  }
}
__device__ __forceinline__ void af_async_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}
__device__ __forceinline__ void af_scale_rows(float* dst, const AfCtx& c, float div) {
  const int lane = threadIdx.x & 31;
  const int w = c.dk4 * 4;
  for (int idx = lane; idx < c.kb * w; idx += 32) {
    const int j = idx / w, cc = idx - j * w;
    dst[j * AF_LDS + cc] = __fdiv_rn(dst[j * AF_LDS + cc], div);
  }
  __syncwarp();
}
__device__ __forceinline__ float af_dot(const float* r, const float* row, int dk4) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four independent chains (FFMA latency 4)
#pragma unroll
  for (int c4 = 0; c4 < AF_DK / 4; ++c4) {
    if (c4 < dk4) {
      const float4 k = *reinterpret_cast<const float4*>(row + c4 * 4);
      a0 = fmaf(r[c4 * 4 + 0], k.x, a0);
      a1 = fmaf(r[c4 * 4 + 1], k.y, a1);
      a2 = fmaf(r[c4 * 4 + 2], k.z, a2);
      a3 = fmaf(r[c4 * 4 + 3], k.w, a3);
    }
  }
  return (a0 + a1) + (a2 + a3);
}
__device__ __forceinline__ void af_axpy(float* r, float s, const float* row, int dk4) {
#pragma unroll
  for (int c4 = 0; c4 < AF_DK / 4; ++c4) {
    if (c4 < dk4) {
      const float4 k = *reinterpret_cast<const float4*>(row + c4 * 4);
      r[c4 * 4 + 0] = fmaf(s, k.x, r[c4 * 4 + 0]);
      r[c4 * 4 + 1] = fmaf(s, k.y, r[c4 * 4 + 1]);
      r[c4 * 4 + 2] = fmaf(s, k.z, r[c4 * 4 + 2]);
      r[c4 * 4 + 3] = fmaf(s, k.w, r[c4 * 4 + 3]);
    }
  }
}
__device__ __forceinline__ void af_row(float* r, const float* row, int dk4) {
#pragma unroll
  for (int c4 = 0; c4 < AF_DK / 4; ++c4) {
    float4 k = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < dk4) k = *reinterpret_cast<const float4*>(row + c4 * 4);
    r[c4 * 4 + 0] = k.x; r[c4 * 4 + 1] = k.y; r[c4 * 4 + 2] = k.z; r[c4 * 4 + 3] = k.w;
  }
}

__global__ void __launch_bounds__(32 * AF_WARPS) attention_fast_fwd_kernel(const AttArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AfCtx c = af_ctx<AF_WARPS>(a);
  if (!c.active) return;
  float* Qs = sm + (size_t)warp * (3 * AF_KMAX * AF_LDS + AF_KMAX * 32);
  float* Ks = Qs + AF_KMAX * AF_LDS;
  float* Vs = Ks + AF_KMAX * AF_LDS;
  float* Ss = Vs + AF_KMAX * AF_LDS;   // [j2][lane] score scratch of this warp
  if (af_vec_ok(a)) {
    af_load_async(a.q, a.ld, c, Qs);
    af_load_async(a.k, a.ld, c, Ks);
    af_load_async(a.v, a.ld, c, Vs);
    af_async_wait();
    af_scale_rows(Qs, c, a.inv_temp_div);
  } else {
    af_load(a.q, a.ld, c, Qs, a.inv_temp_div);
    af_load(a.k, a.ld, c, Ks, 0.f);
    af_load(a.v, a.ld, c, Vs, 0.f);
    __syncwarp();
  }
  const long long node = blockIdx.x;
  for (int j1 = lane; j1 < c.kb; j1 += 32) {
    float q[AF_DK];
    af_row(q, Qs + j1 * AF_LDS, c.dk4);
    float m = -INFINITY;
#pragma unroll 4
    for (int j2 = 0; j2 < c.kb; ++j2) {
      const float sc = af_dot(q, Ks + j2 * AF_LDS, c.dk4);
      Ss[j2 * 32 + lane] = sc;
      m = fmaxf(m, sc);
    }
    float o[AF_DK];
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc) o[cc] = 0.f;
    float sum = 0.f;
    for (int j2 = 0; j2 < c.kb; ++j2) {   // pass 2a: row sum (the reference normalises before dropout and P V)
      const float e = expf(Ss[j2 * 32 + lane] - m);
      Ss[j2 * 32 + lane] = e;
      sum += e;
    }
#pragma unroll 4
    for (int j2 = 0; j2 < c.kb; ++j2) {
      float p = __fdiv_rn(Ss[j2 * 32 + lane], sum);
      if (a.drop_p > 0.f) p *= att_keep_scale(a.seed, node, c.h, j1, j2, a.drop_p);
      af_axpy(o, p, Vs + j2 * AF_LDS, c.dk4);
    }
    float* dst = a.o + (c.r0 + (long long)j1 * c.n) * a.ld + c.h * c.dk;
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc)
      if (cc < c.dk) dst[cc] = o[cc];
  }
}

__global__ void __launch_bounds__(32 * AF_WARPS_B) attention_fast_bwd_kernel(const AttArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AfCtx c = af_ctx<AF_WARPS_B>(a);
  if (!c.active) return;
  float* Qs = sm + (size_t)warp * (4 * AF_KMAX * AF_LDS + 2 * AF_KMAX * 32 + 3 * AF_KMAX);
  float* Ks = Qs + AF_KMAX * AF_LDS;
  float* Vs = Ks + AF_KMAX * AF_LDS;
  float* Gs = Vs + AF_KMAX * AF_LDS;      // dO
  float* Ps = Gs + AF_KMAX * AF_LDS;      // [j2][lane] probabilities (phase A scratch)
  float* Dp = Ps + AF_KMAX * 32;          // [j2][lane] dP            (phase A scratch)
  float* Ms = Dp + AF_KMAX * 32;          // row max
  float* Ls = Ms + AF_KMAX;               // row sum
  float* Ds = Ls + AF_KMAX;               // sum_j2 dP * P
  if (af_vec_ok(a)) {
    af_load_async(a.q, a.ld, c, Qs);
    af_load_async(a.k, a.ld, c, Ks);
    af_load_async(a.v, a.ld, c, Vs);
    af_load_async(a.go, a.ld, c, Gs);
    af_async_wait();
    af_scale_rows(Qs, c, a.inv_temp_div);
  } else {
    af_load(a.q, a.ld, c, Qs, a.inv_temp_div);
    af_load(a.k, a.ld, c, Ks, 0.f);
    af_load(a.v, a.ld, c, Vs, 0.f);
    af_load(a.go, a.ld, c, Gs, 0.f);
    __syncwarp();
  }
  const long long node = blockIdx.x;
  // ---- phase A: lane = query j1 -> row statistics, D, dQ
  for (int j1 = lane; j1 < c.kb; j1 += 32) {
    float q[AF_DK], g[AF_DK];
    af_row(q, Qs + j1 * AF_LDS, c.dk4);
    af_row(g, Gs + j1 * AF_LDS, c.dk4);
    float m = -INFINITY;
#pragma unroll 4
    for (int j2 = 0; j2 < c.kb; ++j2) {
      const float sc = af_dot(q, Ks + j2 * AF_LDS, c.dk4);
      Ps[j2 * 32 + lane] = sc;
      m = fmaxf(m, sc);
    }
    float sum = 0.f;
    for (int j2 = 0; j2 < c.kb; ++j2) {
      const float e = expf(Ps[j2 * 32 + lane] - m);
      Ps[j2 * 32 + lane] = e;
      sum += e;
    }
    float dot = 0.f;
#pragma unroll 4
    for (int j2 = 0; j2 < c.kb; ++j2) {
      float dp = af_dot(g, Vs + j2 * AF_LDS, c.dk4);
      if (a.drop_p > 0.f) dp *= att_keep_scale(a.seed, node, c.h, j1, j2, a.drop_p);
      const float p = __fdiv_rn(Ps[j2 * 32 + lane], sum);
      Ps[j2 * 32 + lane] = p;
      Dp[j2 * 32 + lane] = dp;
      dot = fmaf(dp, p, dot);
    }
    float dq[AF_DK];
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc) dq[cc] = 0.f;
#pragma unroll 4
    for (int j2 = 0; j2 < c.kb; ++j2) {
      const float ds = Ps[j2 * 32 + lane] * (Dp[j2 * 32 + lane] - dot);
      af_axpy(dq, ds, Ks + j2 * AF_LDS, c.dk4);
    }
    Ms[j1] = m;
    Ls[j1] = sum;
    Ds[j1] = dot;
    float* dst = a.gq + (c.r0 + (long long)j1 * c.n) * a.ld + c.h * c.dk;
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc)
      if (cc < c.dk) dst[cc] = __fdiv_rn(dq[cc], a.inv_temp_div);
  }
  __syncwarp();
  // ---- phase B: lane = key j2 -> dK, dV (P and dS recomputed column-wise from Q, K and the row statistics)
  for (int j2 = lane; j2 < c.kb; j2 += 32) {
    float kr[AF_DK], vr[AF_DK], dk_[AF_DK], dv_[AF_DK];
    af_row(kr, Ks + j2 * AF_LDS, c.dk4);
    af_row(vr, Vs + j2 * AF_LDS, c.dk4);
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc) { dk_[cc] = 0.f; dv_[cc] = 0.f; }
#pragma unroll 4
    for (int j1 = 0; j1 < c.kb; ++j1) {
      const float sc = af_dot(kr, Qs + j1 * AF_LDS, c.dk4);
      float dp = af_dot(vr, Gs + j1 * AF_LDS, c.dk4);
      const float p = __fdiv_rn(expf(sc - Ms[j1]), Ls[j1]);
      float pd = p;
      if (a.drop_p > 0.f) {
        const float ks = att_keep_scale(a.seed, node, c.h, j1, j2, a.drop_p);
        pd *= ks;
        dp *= ks;
      }
      const float ds = p * (dp - Ds[j1]);
      af_axpy(dv_, pd, Gs + j1 * AF_LDS, c.dk4);
      af_axpy(dk_, ds, Qs + j1 * AF_LDS, c.dk4);
    }
    float* dstk = a.gk + (c.r0 + (long long)j2 * c.n) * a.ld + c.h * c.dk;
    float* dstv = a.gv + (c.r0 + (long long)j2 * c.n) * a.ld + c.h * c.dk;
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc) {
      if (cc < c.dk) {
        dstk[cc] = dk_[cc];
        dstv[cc] = dv_[cc];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// v2 (16-byte aligned head slices): only the two tiles a phase broadcasts live in shared memory - K, V while a lane is
// a query, Q, dO while a lane is a key - and the lane's own rows come straight from global memory into registers.
// Shared memory per warp drops from 22.4 / 33.8 KB to 16.5 / 21.6 KB, i.e. 12 / 10 instead of 8 / 6 resident warps per
// SM for a kernel that is bound by the latency of its dependent FMA chains (profiles/r1h_launches_step.csv: 4.6 ms for
// 2 GB of traffic and 18 GFLOP).
__device__ __forceinline__ void af_row_global(float* r, const float* __restrict__ src, int dk4) {
#pragma unroll
  for (int c4 = 0; c4 < AF_DK / 4; ++c4) {
    float4 k = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < dk4) k = ldg4(src + c4 * 4);
    r[c4 * 4 + 0] = k.x; r[c4 * 4 + 1] = k.y; r[c4 * 4 + 2] = k.z; r[c4 * 4 + 3] = k.w;
  }
}
#define AF2_FWD_FLOATS (2 * AF_KMAX * AF_LDS + AF_KMAX * 32)
#define AF2_BWD_FLOATS (2 * AF_KMAX * AF_LDS + 2 * AF_KMAX * 32 + 3 * AF_KMAX)

// DK4 > 0: head width fixed at compile time (dk = 4 * DK4) - no per-FMA predicates; DK4 = 0: runtime width.
// Loop bodies are kept small on purpose: the first v2 build unrolled the (j1, j2) loops four times and spent 36 % of
// its issue slots waiting for instruction fetch (21 KB loop body, profiles/r1j_att_bwd_ncu.csv).
template <int DK4>
__global__ void __launch_bounds__(32 * AF_WARPS) attention_fast2_fwd_kernel(const AttArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AfCtx c = af_ctx<AF_WARPS>(a);
  if (!c.active) return;
  const int dk4 = DK4 > 0 ? DK4 : c.dk4;
  float* Ks = sm + (size_t)warp * AF2_FWD_FLOATS;
  float* Vs = Ks + AF_KMAX * AF_LDS;
  float* Ss = Vs + AF_KMAX * AF_LDS;   // [j2][lane] score scratch of this warp
  af_load_async(a.k, a.ld, c, Ks);
  af_load_async(a.v, a.ld, c, Vs);
  const long long node = blockIdx.x;
  float q[AF_DK];
  if (lane < c.kb) af_row_global(q, a.q + (c.r0 + (long long)lane * c.n) * a.ld + c.h * c.dk, dk4);
  af_async_wait();   // every lane: its own cp.async group, then the warp barrier that publishes the tiles
  for (int j1 = lane; j1 < c.kb; j1 += 32) {
    if (j1 != lane) af_row_global(q, a.q + (c.r0 + (long long)j1 * c.n) * a.ld + c.h * c.dk, dk4);
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc) q[cc] = __fdiv_rn(q[cc], a.inv_temp_div);
    float m = -INFINITY;
#pragma unroll 2
    for (int j2 = 0; j2 < c.kb; ++j2) {
      const float sc = af_dot(q, Ks + j2 * AF_LDS, dk4);
      Ss[j2 * 32 + lane] = sc;
      m = fmaxf(m, sc);
    }
    float o[AF_DK];
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc) o[cc] = 0.f;
    float sum = 0.f;
    for (int j2 = 0; j2 < c.kb; ++j2) {   // pass 2a: row sum (the reference normalises before dropout and P V)
      const float e = expf(Ss[j2 * 32 + lane] - m);
      Ss[j2 * 32 + lane] = e;
      sum += e;
    }
#pragma unroll 2
    for (int j2 = 0; j2 < c.kb; ++j2) {
      float p = __fdiv_rn(Ss[j2 * 32 + lane], sum);
      if (a.drop_p > 0.f) p *= att_keep_scale(a.seed, node, c.h, j1, j2, a.drop_p);
      af_axpy(o, p, Vs + j2 * AF_LDS, dk4);
    }
    float* dst = a.o + (c.r0 + (long long)j1 * c.n) * a.ld + c.h * c.dk;
#pragma unroll
    for (int c4 = 0; c4 < AF_DK / 4; ++c4)
      if (c4 < dk4)
        *reinterpret_cast<float4*>(dst + c4 * 4) = make_float4(o[c4 * 4], o[c4 * 4 + 1], o[c4 * 4 + 2], o[c4 * 4 + 3]);
  }
}

template <int DK4>
__global__ void __launch_bounds__(32 * AF_WARPS_B, 4) attention_fast2_bwd_kernel(const AttArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const AfCtx c = af_ctx<AF_WARPS_B>(a);
  if (!c.active) return;
  const int dk4 = DK4 > 0 ? DK4 : c.dk4;
  float* T0 = sm + (size_t)warp * AF2_BWD_FLOATS;   // K, then Q
  float* T1 = T0 + AF_KMAX * AF_LDS;                // V, then dO
  float* Ps = T1 + AF_KMAX * AF_LDS;                // [j2][lane] probabilities (phase A scratch)
  float* Dp = Ps + AF_KMAX * 32;                    // [j2][lane] dP            (phase A scratch)
  float* Ms = Dp + AF_KMAX * 32;                    // row max
  float* Ls = Ms + AF_KMAX;                         // row sum
  float* Ds = Ls + AF_KMAX;                         // sum_j2 dP * P
  af_load_async(a.k, a.ld, c, T0);
  af_load_async(a.v, a.ld, c, T1);
  const long long node = blockIdx.x;
  // ---- phase A: lane = query j1 -> row statistics, D, dQ   (K, V tiles in shared memory; q, dO rows in registers)
  {
    float q[AF_DK], g[AF_DK];
    if (lane < c.kb) {
      const long long roff0 = (c.r0 + (long long)lane * c.n) * a.ld + c.h * c.dk;
      af_row_global(q, a.q + roff0, dk4);
      af_row_global(g, a.go + roff0, dk4);
    }
    af_async_wait();   // every lane: its own cp.async group, then the warp barrier that publishes the tiles
  for (int j1 = lane; j1 < c.kb; j1 += 32) {
    const long long roff = (c.r0 + (long long)j1 * c.n) * a.ld + c.h * c.dk;
    if (j1 != lane) {
      af_row_global(q, a.q + roff, dk4);
      af_row_global(g, a.go + roff, dk4);
    }
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc) q[cc] = __fdiv_rn(q[cc], a.inv_temp_div);
    float m = -INFINITY;
#pragma unroll 2
    for (int j2 = 0; j2 < c.kb; ++j2) {
      const float sc = af_dot(q, T0 + j2 * AF_LDS, dk4);
      Ps[j2 * 32 + lane] = sc;
      m = fmaxf(m, sc);
    }
    float sum = 0.f;
    for (int j2 = 0; j2 < c.kb; ++j2) {
      const float e = expf(Ps[j2 * 32 + lane] - m);
      Ps[j2 * 32 + lane] = e;
      sum += e;
    }
    float dot = 0.f;
#pragma unroll 2
    for (int j2 = 0; j2 < c.kb; ++j2) {
      float dp = af_dot(g, T1 + j2 * AF_LDS, dk4);
      if (a.drop_p > 0.f) dp *= att_keep_scale(a.seed, node, c.h, j1, j2, a.drop_p);
      const float p = __fdiv_rn(Ps[j2 * 32 + lane], sum);
      Ps[j2 * 32 + lane] = p;
      Dp[j2 * 32 + lane] = dp;
      dot = fmaf(dp, p, dot);
    }
    float dq[AF_DK];
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc) dq[cc] = 0.f;
#pragma unroll 2
    for (int j2 = 0; j2 < c.kb; ++j2) {
      const float ds = Ps[j2 * 32 + lane] * (Dp[j2 * 32 + lane] - dot);
      af_axpy(dq, ds, T0 + j2 * AF_LDS, dk4);
    }
    Ms[j1] = m;
    Ls[j1] = sum;
    Ds[j1] = dot;
    float* dst = a.gq + roff;
#pragma unroll
    for (int c4 = 0; c4 < AF_DK / 4; ++c4)
      if (c4 < dk4)
        *reinterpret_cast<float4*>(dst + c4 * 4) =
            make_float4(__fdiv_rn(dq[c4 * 4], a.inv_temp_div), __fdiv_rn(dq[c4 * 4 + 1], a.inv_temp_div),
                        __fdiv_rn(dq[c4 * 4 + 2], a.inv_temp_div), __fdiv_rn(dq[c4 * 4 + 3], a.inv_temp_div));
  }
  }
  __syncwarp();
  // ---- phase B: lane = key j2 -> dK, dV (P and dS recomputed column-wise from Q, K and the row statistics).
  // The two tiles are refilled with Q and dO; the lane's own K / V rows come from global memory.
  af_load_async(a.q, a.ld, c, T0);
  af_load_async(a.go, a.ld, c, T1);
  float kr[AF_DK], vr[AF_DK];
  if (lane < c.kb) {
    const long long roff0 = (c.r0 + (long long)lane * c.n) * a.ld + c.h * c.dk;
    af_row_global(kr, a.k + roff0, dk4);
    af_row_global(vr, a.v + roff0, dk4);
  }
  af_async_wait();
  af_scale_rows(T0, c, a.inv_temp_div);   // all lanes: q / temperature, as the forward used it
  for (int j2 = lane; j2 < c.kb; j2 += 32) {
    float dk_[AF_DK], dv_[AF_DK];
    const long long roff = (c.r0 + (long long)j2 * c.n) * a.ld + c.h * c.dk;
    if (j2 != lane) {
      af_row_global(kr, a.k + roff, dk4);
      af_row_global(vr, a.v + roff, dk4);
    }
#pragma unroll
    for (int cc = 0; cc < AF_DK; ++cc) { dk_[cc] = 0.f; dv_[cc] = 0.f; }
#pragma unroll 1
    for (int j1 = 0; j1 < c.kb; ++j1) {
      const float sc = af_dot(kr, T0 + j1 * AF_LDS, dk4);
      float dp = af_dot(vr, T1 + j1 * AF_LDS, dk4);
      const float p = __fdiv_rn(expf(sc - Ms[j1]), Ls[j1]);
      float pd = p;
      if (a.drop_p > 0.f) {
        const float ks = att_keep_scale(a.seed, node, c.h, j1, j2, a.drop_p);
        pd *= ks;
        dp *= ks;
      }
      const float ds = p * (dp - Ds[j1]);
      af_axpy(dv_, pd, T1 + j1 * AF_LDS, dk4);
      af_axpy(dk_, ds, T0 + j1 * AF_LDS, dk4);
    }
    float* dstk = a.gk + roff;
    float* dstv = a.gv + roff;
#pragma unroll
    for (int c4 = 0; c4 < AF_DK / 4; ++c4) {
      if (c4 < dk4) {
        *reinterpret_cast<float4*>(dstk + c4 * 4) = make_float4(dk_[c4 * 4], dk_[c4 * 4 + 1], dk_[c4 * 4 + 2], dk_[c4 * 4 + 3]);
        *reinterpret_cast<float4*>(dstv + c4 * 4) = make_float4(dv_[c4 * 4], dv_[c4 * 4 + 1], dv_[c4 * 4 + 2], dv_[c4 * 4 + 3]);
      }
    }
  }
}

static bool af_vec_host(const AttArgs& a) {
  auto al = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
  return (a.dk & 3) == 0 && (a.ld & 3) == 0 && al(a.q) && al(a.k) && al(a.v) && al(a.o);
}
int sb_attention_fast_fwd_launch(const AttArgs& a, int kmax, cudaStream_t st) {
  if (kmax > AF_KMAX || a.dk > AF_DK) return SB_ERR_UNSUPPORTED;
  dim3 grid((unsigned)a.N, (unsigned)((a.n_head + AF_WARPS - 1) / AF_WARPS));
  if (af_vec_host(a)) {
    const size_t smem2 = (size_t)AF_WARPS * AF2_FWD_FLOATS * sizeof(float);
    static bool configured2 = false;
    if (!configured2) {
      SB_CUDA(cudaFuncSetAttribute(attention_fast2_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      SB_CUDA(cudaFuncSetAttribute(attention_fast2_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      configured2 = true;
    }
    if (a.dk == 32) attention_fast2_fwd_kernel<8><<<grid, 32 * AF_WARPS, smem2, st>>>(a);
    else attention_fast2_fwd_kernel<0><<<grid, 32 * AF_WARPS, smem2, st>>>(a);
    SB_CHECK_LAUNCH("sb_attention_fwd(fast2)");
    return SB_OK;
  }
  const size_t smem = (size_t)AF_WARPS * (3 * AF_KMAX * AF_LDS + AF_KMAX * 32) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SB_CUDA(cudaFuncSetAttribute(attention_fast_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  attention_fast_fwd_kernel<<<grid, 32 * AF_WARPS, smem, st>>>(a);
  SB_CHECK_LAUNCH("sb_attention_fwd(fast)");
  return SB_OK;
}
int sb_attention_fast_bwd_launch(const AttArgs& a, int kmax, cudaStream_t st) {
  if (kmax > AF_KMAX || a.dk > AF_DK) return SB_ERR_UNSUPPORTED;
  dim3 grid((unsigned)a.N, (unsigned)((a.n_head + AF_WARPS_B - 1) / AF_WARPS_B));
  auto al = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
  if ((a.dk & 3) == 0 && (a.ld & 3) == 0 && al(a.q) && al(a.k) && al(a.v) && al(a.go) && al(a.gq) && al(a.gk) && al(a.gv)) {
    const size_t smem2 = (size_t)AF_WARPS_B * AF2_BWD_FLOATS * sizeof(float);
    static bool configured2 = false;
    if (!configured2) {
      SB_CUDA(cudaFuncSetAttribute(attention_fast2_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      SB_CUDA(cudaFuncSetAttribute(attention_fast2_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      configured2 = true;
    }
    if (a.dk == 32) attention_fast2_bwd_kernel<8><<<grid, 32 * AF_WARPS_B, smem2, st>>>(a);
    else attention_fast2_bwd_kernel<0><<<grid, 32 * AF_WARPS_B, smem2, st>>>(a);
    SB_CHECK_LAUNCH("sb_attention_bwd(fast2)");
    return SB_OK;
  }
  const size_t smem = (size_t)AF_WARPS_B * (4 * AF_KMAX * AF_LDS + 2 * AF_KMAX * 32 + 3 * AF_KMAX) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SB_CUDA(cudaFuncSetAttribute(attention_fast_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  attention_fast_bwd_kernel<<<grid, 32 * AF_WARPS_B, smem, st>>>(a);
  SB_CHECK_LAUNCH("sb_attention_bwd(fast)");
  return SB_OK;
}

// K1 — the GIN neighbourhood aggregate of SignNet's phi on the ragged slot-row layout:
//     out[s, (b,j,i), :] = (1 + eps) * x[s, (b,j,i), :] + sum_{(u -> i) in E_b} x[s, (b,j,u), :]
// Replaces torch_geometric GINConv(Identity(), train_eps=True) applied on node dim -2 of a [k,N,d] tensor
// (Alchemy/sign_net/model_utils/masked_layers.py:70,75) and dgl GINConv(..., 'sum') (GraphPrediction/layers/gnns.py:90-98),
// which gather [k,E,d] messages and scatter them back with atomics.
//
// B200 design: every (sign, graph, slot-chunk) is one contiguous [rows, ld] tile (<= 32 KB) that is closed under the
// neighbourhood relation, so a persistent CTA per SM streams tiles HBM -> shared memory with the TMA engine
// (cp.async.bulk + mbarrier full/empty ring, one producer warp), consumer warps do the neighbour sums out of shared
// memory in a fixed CSR order (no atomics, deterministic, bit-identical to the CPU reference's edge-order
// accumulation) and stream the result back with 128-bit coalesced stores.  HBM traffic = read once + write once.
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define AGG_STAGES 5
#define AGG_TILE_BYTES 32768
#define AGG_CONSUMER_WARPS 16
#define AGG_THREADS (32 * (1 + AGG_CONSUMER_WARPS))

struct AggArgs {
  const float* x;
  float* out;
  const float* res;     // optional: out = res + aggregate   (backward: gradient arriving through the residual)
  const float* dotx;    // optional: accumulate sum(x * dotx) into dot_out (d eps)
  double* dot_out;
  const float* eps;     // device scalar or null (eps = 0)
  const int32_t* graph_ptr;
  const int32_t* unit_ptr;
  const int64_t* row_ptr;
  const int32_t* nbr_ptr;
  const int32_t* nbr_idx;
  int64_t R;
  int B, k, masked, S, ld, tile_rows;
};

struct UnitDesc {
  long long row0;  // first row of the tile (sign offset included)
  int n;           // nodes of the graph
  int rows;        // rows in this tile (= slots * n)
  int node0;       // first global node id of the graph
  int pad;
};

__device__ __forceinline__ int upper_graph_i32(const int32_t* __restrict__ ptr, int B, int v) {
  // largest b in [0,B) with ptr[b] <= v   (ptr non-decreasing, ptr[B] > v)
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid) <= v) lo = mid; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_graph_i64(const int64_t* __restrict__ ptr, int B, long long v) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid) <= v) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(AGG_THREADS, 1) gin_agg_tma_kernel(const AggArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)AGG_STAGES * AGG_TILE_BYTES);
  uint64_t* empty = full + AGG_STAGES;
  UnitDesc* desc = reinterpret_cast<UnitDesc*>(empty + AGG_STAGES);
  __shared__ double s_dot[AGG_CONSUMER_WARPS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int U = __ldg(a.unit_ptr + a.B);
  const long long total = (long long)U * a.S;

  if (threadIdx.x == 0) {
    for (int s = 0; s < AGG_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], AGG_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: one lane drives the TMA engine
    if (lane == 0) {
      int it = 0;
      for (long long u = blockIdx.x; u < total; u += gridDim.x, ++it) {
        const int stage = it % AGG_STAGES;
        const uint32_t phase = (uint32_t)(it / AGG_STAGES) & 1u;
        mbar_wait(&empty[stage], phase ^ 1u);
        const int s = (int)(u / U), uu = (int)(u % U);
        const int b = upper_graph_i32(a.unit_ptr, a.B, uu);
        const int c = uu - __ldg(a.unit_ptr + b);
        const int node0 = __ldg(a.graph_ptr + b);
        const int n = __ldg(a.graph_ptr + b + 1) - node0;
        const int kb = a.masked ? (n < a.k ? n : a.k) : a.k;
        int G = a.tile_rows / n;
        if (G < 1) G = 1;
        const int j0 = c * G;
        const int ns = (kb - j0 < G) ? (kb - j0) : G;
        UnitDesc d;
        d.row0 = (long long)s * a.R + __ldg(a.row_ptr + b) + (long long)j0 * n;
        d.n = n;
        d.rows = ns * n;
        d.node0 = node0;
        d.pad = 0;
        desc[stage] = d;
        const uint32_t bytes = (uint32_t)d.rows * (uint32_t)a.ld * 4u;
        mbar_arrive_expect_tx(&full[stage], bytes);
        bulk_g2s(tiles + (size_t)stage * (AGG_TILE_BYTES / 4), a.x + d.row0 * a.ld, bytes, &full[stage]);
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers: neighbour sums out of smem
    const int cw = warp - 1;
    const int ctid = threadIdx.x - 32;
    const float one_eps = __fadd_rn(1.0f, a.eps ? __ldg(a.eps) : 0.0f);
    const int ld = a.ld, ld4 = ld >> 2;
    double dot = 0.0;
    int it = 0;
    for (long long u = blockIdx.x; u < total; u += gridDim.x, ++it) {
      const int stage = it % AGG_STAGES;
      const uint32_t phase = (uint32_t)(it / AGG_STAGES) & 1u;
      mbar_wait(&full[stage], phase);
      const UnitDesc d = desc[stage];
      const float* tile = tiles + (size_t)stage * (AGG_TILE_BYTES / 4);
      // work items = (row, float4 column) pairs, flattened over all consumer threads
      const int items = d.rows * ld4;
#pragma unroll 2
      for (int i = ctid; i < items; i += 32 * AGG_CONSUMER_WARPS) {
        const int r = i / ld4;
        const int c4 = i - r * ld4;
        const int slot = r / d.n;
        const int li = r - slot * d.n;
        const int beg = __ldg(a.nbr_ptr + d.node0 + li), end = __ldg(a.nbr_ptr + d.node0 + li + 1);
        const float* slot_tile = tile + (size_t)slot * d.n * ld + c4 * 4;
        const long long g = (d.row0 + r) * (long long)ld + c4 * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = beg; e < end; ++e) {
          const int lj = __ldg(a.nbr_idx + e) - d.node0;
          const float4 v = *reinterpret_cast<const float4*>(slot_tile + (size_t)lj * ld);
          acc.x = __fadd_rn(acc.x, v.x);
          acc.y = __fadd_rn(acc.y, v.y);
          acc.z = __fadd_rn(acc.z, v.z);
          acc.w = __fadd_rn(acc.w, v.w);
        }
        const float4 self = *reinterpret_cast<const float4*>(slot_tile + (size_t)li * ld);
        acc.x = __fadd_rn(acc.x, __fmul_rn(one_eps, self.x));
        acc.y = __fadd_rn(acc.y, __fmul_rn(one_eps, self.y));
        acc.z = __fadd_rn(acc.z, __fmul_rn(one_eps, self.z));
        acc.w = __fadd_rn(acc.w, __fmul_rn(one_eps, self.w));
        if (a.res) {  // may alias `out`: plain (coherent) load, read before the store below
          const float4 q = *reinterpret_cast<const float4*>(a.res + g);
          acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
        }
        if (a.dotx) {
          const float4 t = ldg4(a.dotx + g);
          dot += (double)(self.x * t.x + self.y * t.y + self.z * t.z + self.w * t.w);
        }
        stg4_stream(a.out + g, acc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
    if (a.dotx) {
      dot = warp_sum_d(dot);
      if (lane == 0) s_dot[cw] = dot;
    }
  }
  if (a.dotx) {
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < AGG_CONSUMER_WARPS; ++w) t += s_dot[w];
      atomicAdd(a.dot_out, t);
    }
  }
}

// Generic fallback: one thread per (row, column), neighbours read through L1/L2.  Used for d_in = 1 (layer 0), for row
// strides that are not a multiple of 4 floats and for graphs too large for one shared-memory tile.
__global__ void gin_agg_generic_kernel(const AggArgs a) {
  const long long rows_total = (long long)a.S * a.R;
  const long long total = rows_total * a.ld;
  const float one_eps = __fadd_rn(1.0f, a.eps ? __ldg(a.eps) : 0.0f);
  double dot = 0.0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / a.ld;
    const int c = (int)(t - row * a.ld);
    const int s = (int)(row / a.R);
    const long long r = row - (long long)s * a.R;
    const int b = upper_graph_i64(a.row_ptr, a.B, r);
    const int node0 = __ldg(a.graph_ptr + b);
    const int n = __ldg(a.graph_ptr + b + 1) - node0;
    const long long off = r - __ldg(a.row_ptr + b);
    const int slot = (int)(off / n);
    const int li = (int)(off - (long long)slot * n);
    const long long slot_row0 = (long long)s * a.R + __ldg(a.row_ptr + b) + (long long)slot * n;
    const int beg = __ldg(a.nbr_ptr + node0 + li), end = __ldg(a.nbr_ptr + node0 + li + 1);
    float acc = 0.f;
    for (int e = beg; e < end; ++e) {
      const int lj = __ldg(a.nbr_idx + e) - node0;
      acc = __fadd_rn(acc, __ldg(a.x + (slot_row0 + lj) * a.ld + c));
    }
    const float self = __ldg(a.x + t);
    acc = __fadd_rn(acc, __fmul_rn(one_eps, self));
    if (a.res) acc += a.res[t];
    if (a.dotx) dot += (double)(self * __ldg(a.dotx + t));
    a.out[t] = acc;
  }
  if (a.dotx) {
    dot = warp_sum_d(dot);
    if ((threadIdx.x & 31) == 0 && dot != 0.0) atomicAdd(a.dot_out, dot);
  }
}

extern "C" int sb_gin_agg(const float* x, float* out, const float* res, const float* dotx, double* dot_out,
                          const float* eps, const int32_t* graph_ptr, const int32_t* unit_ptr,
                          const int64_t* row_ptr, const int32_t* nbr_ptr, const int32_t* nbr_idx, int64_t R,
                          int32_t B, int32_t k, int32_t masked, int32_t S, int32_t ld, int32_t tile_rows,
                          int32_t force_generic, void* stream) {
  SB_CHECK_ARG(R >= 0 && B >= 0 && S >= 1 && ld >= 1, "sb_gin_agg: bad sizes");
  SB_CHECK_ARG((dotx == nullptr) == (dot_out == nullptr), "sb_gin_agg: dotx and dot_out must be given together");
  SB_CHECK_ARG(x != out, "sb_gin_agg: x and out must not alias");
  if (R == 0 || B == 0) return SB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  AggArgs a;
  a.x = x; a.out = out; a.res = res; a.dotx = dotx; a.dot_out = dot_out; a.eps = eps;
  a.graph_ptr = graph_ptr; a.unit_ptr = unit_ptr; a.row_ptr = row_ptr; a.nbr_ptr = nbr_ptr; a.nbr_idx = nbr_idx;
  a.R = R; a.B = B; a.k = k; a.masked = masked; a.S = S; a.ld = ld; a.tile_rows = tile_rows;
  const bool tma_ok = !force_generic && (ld % 4 == 0) && tile_rows >= 1 &&
                      (int64_t)tile_rows * ld * 4 <= AGG_TILE_BYTES && ((uintptr_t)x % 16 == 0) &&
                      ((uintptr_t)out % 16 == 0) && (!res || (uintptr_t)res % 16 == 0) &&
                      (!dotx || (uintptr_t)dotx % 16 == 0);
  if (tma_ok) {
    static bool attr_set = false;
    const size_t smem = (size_t)AGG_STAGES * AGG_TILE_BYTES + 2 * AGG_STAGES * sizeof(uint64_t) +
                        AGG_STAGES * sizeof(UnitDesc);
    if (!attr_set) {
      SB_CUDA(cudaFuncSetAttribute(gin_agg_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    gin_agg_tma_kernel<<<sb_num_sms(), AGG_THREADS, smem, st>>>(a);
    SB_CHECK_LAUNCH("sb_gin_agg(tma)");
  } else {
    const long long total = (long long)S * R * ld;
    long long blocks = sb_ceil_div(total, 256);
    const long long cap = (long long)sb_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    gin_agg_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    SB_CHECK_LAUNCH("sb_gin_agg(generic)");
  }
  return SB_OK;
}

extern "C" int sb_gin_agg_tile_rows(int32_t ld) {
  if (ld <= 0 || ld % 4 != 0) return 0;
  return AGG_TILE_BYTES / (ld * 4);
}

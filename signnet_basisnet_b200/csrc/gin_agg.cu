// K1 — the GIN neighbourhood aggregate of SignNet's phi on the ragged slot-row layout:
//     out[s, (b,j,i), :] = (1 + eps) * x[s, (b,j,i), :] + sum_{(u -> i) in E_b} x[s, (b,j,u), :]
// Replaces torch_geometric GINConv(Identity(), train_eps=True) applied on node dim -2 of a [k,N,d] tensor
// (Alchemy/sign_net/model_utils/masked_layers.py:70,75) and dgl GINConv(..., 'sum') (GraphPrediction/layers/gnns.py:90-98),
// which gather [k,E,d] messages and scatter them back with atomics.
//
// B200 design: every (sign, graph, slot-chunk) is one contiguous [rows, ld] tile (<= 32 KB) that is closed under the
// neighbourhood relation, so a persistent CTA per SM streams tiles HBM -> shared memory with the TMA engine
// (cp.async.bulk + mbarrier full/empty ring).  The producer warp finds a tile with one coalesced load from the unit
// table (sb_agg_unit_desc), stages the graph's packed neighbour words (4 local ids per node in one 32-bit word,
// sb_pack_neighbours) next to the tile and - in the backward - also the residual-gradient and d-eps operand tiles, so
// the consumer warps touch global memory only for their 128-bit coalesced streaming stores: no dependent global load
// sits on the critical path (the first version spent 60 % of its issue slots stalled on exactly those, see
// profiles/r1b_gin_agg_ncu.csv).  Consumers own whole rows (a warp covers 32/LPR rows per step, LPR = lanes per row):
// one LDS.32 yields the row's neighbour list, four independent LDS.128 fetch the neighbour rows, the adds are
// predicated on the degree and run in CSR (= edge id) order - no atomics, deterministic, bit-identical to the CPU
// reference's edge-order accumulation.  HBM traffic = every operand read once + the result written once.
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define AGG_TILE_BYTES 32768
#define AGG_CONSUMER_WARPS 16
#define AGG_THREADS (32 * (1 + AGG_CONSUMER_WARPS))
#define AGG_MAX_STAGES 6
#define AGG_NB_MAXN 256         // nodes per graph whose neighbour words are staged (larger graphs: global CSR path)
#define AGG_SMEM_LIMIT (227 * 1024)
#define AGG_ZROW_FLOATS 1024     // zero row for missing neighbours: the TMA path takes ld <= 1024 floats
#define NB_SLOW 0xFEu           // byte 3 of a neighbour word: degree > 4 or a local id > 253 -> walk the global CSR
#define NB_NONE 0xFFu           // empty slot of a neighbour word

struct AggArgs {
  const float* x;
  float* out;
  const float* res;     // optional: out = res + aggregate   (backward: gradient arriving through the residual)
  const float* dotx;    // optional: accumulate sum(x * dotx) into dot_out (d eps)
  double* dot_out;
  const float* eps;     // device scalar or null (eps = 0)
  const int32_t* graph_ptr;
  const int32_t* unit_ptr;
  const int32_t* unit_desc;   // [U][12] records written by sb_agg_unit_desc (bookkeeping.cu)
  const uint32_t* nbr_pack;   // [N] packed neighbour words written by sb_pack_neighbours (bookkeeping.cu)
  const int64_t* row_ptr;
  const int32_t* nbr_ptr;
  const int32_t* nbr_idx;
  int64_t R;
  int B, k, masked, S, ld, tile_rows;
  int stages;           // TMA ring depth (host: what fits next to the 1..3 operand tiles per stage)
};

struct UnitDesc {
  long long row0;   // first row of the tile (sign offset included)
  int n;            // nodes of the graph
  int rows;         // rows in this tile (= slots * n)
  int node0;        // first global node id of the graph
  int nb_local;     // 1: neighbour words staged in shared memory, 0: read the CSR from global memory
  unsigned magic;   // ceil(2^32 / n): r / n == (r * magic) >> 32 for r < 2^16
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

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// One output float4 of one row: neighbours (missing ones point at the zero row: x + 0.0f is exact), self term, optional
// residual / d-eps operands from their staged tiles.
template <bool HAS_RES, bool HAS_DOT>
__device__ __forceinline__ void agg_cell(const float* p0, const float* p1, const float* p2, const float* p3,
                                         const float* self_row, int res_off, int dot_off, float one_eps, float* out_row,
                                         int col, double& dot) {
  const float4 v0 = lds4(p0 + col), v1 = lds4(p1 + col), v2 = lds4(p2 + col), v3 = lds4(p3 + col);
  const float4 self = lds4(self_row + col);
  float4 acc = add4(add4(add4(add4(make_float4(0.f, 0.f, 0.f, 0.f), v0), v1), v2), v3);
  acc.x = __fadd_rn(acc.x, __fmul_rn(one_eps, self.x));
  acc.y = __fadd_rn(acc.y, __fmul_rn(one_eps, self.y));
  acc.z = __fadd_rn(acc.z, __fmul_rn(one_eps, self.z));
  acc.w = __fadd_rn(acc.w, __fmul_rn(one_eps, self.w));
  if (HAS_RES) {  // the residual tile was staged before any row of this tile is overwritten (res may alias out)
    const float4 q = lds4(self_row + res_off + col);
    acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
  }
  if (HAS_DOT) {
    const float4 t = lds4(self_row + dot_off + col);
    dot += (double)(self.x * t.x + self.y * t.y + self.z * t.z + self.w * t.w);
  }
  stg4_stream(out_row + col, acc);
}

// shared memory: [stages][T operand tiles of AGG_TILE_BYTES] | full[] empty[] | desc[] | nb[stages][AGG_NB_MAXN] | zero row
template <int LPR, bool HAS_RES, bool HAS_DOT>  // LPR lanes per row: 32, 16, 8, 4 (ld/4 <= LPR, or LPR == 32 + column loop)
__global__ void __launch_bounds__(AGG_THREADS, 1) gin_agg_tma_kernel(const AggArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int T = 1 + (HAS_RES ? 1 : 0) + (HAS_DOT ? 1 : 0);
  const int stages = a.stages;
  constexpr size_t stage_bytes = (size_t)T * AGG_TILE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + stages * stage_bytes);
  uint64_t* empty = full + AGG_MAX_STAGES;
  UnitDesc* desc = reinterpret_cast<UnitDesc*>(empty + AGG_MAX_STAGES);
  uint32_t* nbw = reinterpret_cast<uint32_t*>(desc + AGG_MAX_STAGES);
  float* zrow = reinterpret_cast<float*>(nbw + AGG_MAX_STAGES * AGG_NB_MAXN);   // AGG_ZROW_FLOATS zeros
  __shared__ double s_dot[AGG_CONSUMER_WARPS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int U = __ldg(a.unit_ptr + a.B);
  const long long total = (long long)U * a.S;

  for (int i = threadIdx.x; i < AGG_ZROW_FLOATS; i += blockDim.x) zrow[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], AGG_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------- producer warp: unit record (prefetched one tile ahead), neighbour words, TMA issue
    // lanes 0..5 hold the first words of a unit record; the record of the NEXT tile is requested before this tile's
    // neighbour-word round trip, so the exposed latency per tile is one round of coalesced 4-byte loads.
    int stage = 0;
    uint32_t phase = 0;
    long long u = blockIdx.x;
    int w_next = 0;
    if (u < total && lane < 6) w_next = __ldg(a.unit_desc + (u % U) * 12 + lane);
    for (; u < total; u += gridDim.x) {
      const int w = w_next;
      const long long un = u + gridDim.x;
      if (un < total && lane < 6) w_next = __ldg(a.unit_desc + (un % U) * 12 + lane);
      const int s = (int)(u / U);
      const unsigned rlo = (unsigned)__shfl_sync(0xffffffffu, w, 0), rhi = (unsigned)__shfl_sync(0xffffffffu, w, 1);
      const int n = __shfl_sync(0xffffffffu, w, 2), rows = __shfl_sync(0xffffffffu, w, 3);
      const int node0 = __shfl_sync(0xffffffffu, w, 4);
      const unsigned magic = (unsigned)__shfl_sync(0xffffffffu, w, 5);
      const long long row0 = (long long)s * a.R + (long long)(((unsigned long long)rhi << 32) | rlo);
      const int local = (n <= AGG_NB_MAXN) ? 1 : 0;
      uint32_t nv[AGG_NB_MAXN / 32];
      if (local) {
#pragma unroll
        for (int q = 0; q < AGG_NB_MAXN / 32; ++q) {
          const int i = lane + 32 * q;
          nv[q] = (i < n) ? __ldg(a.nbr_pack + node0 + i) : 0u;
        }
      }
      mbar_wait(&empty[stage], phase ^ 1u);
      if (local) {
        uint32_t* dst = nbw + stage * AGG_NB_MAXN;
#pragma unroll
        for (int q = 0; q < AGG_NB_MAXN / 32; ++q) {
          const int i = lane + 32 * q;
          if (i < n) dst[i] = nv[q];
        }
      }
      __syncwarp();
      if (lane == 0) {
        UnitDesc d;
        d.row0 = row0; d.n = n; d.rows = rows; d.node0 = node0; d.nb_local = local; d.magic = magic; d.pad = 0;
        desc[stage] = d;
        const uint32_t bytes = (uint32_t)rows * (uint32_t)a.ld * 4u;
        float* t0 = reinterpret_cast<float*>(smem_raw + stage * stage_bytes);
        mbar_arrive_expect_tx(&full[stage], bytes * (uint32_t)T);   // release: publishes desc + words to consumers
        bulk_g2s(t0, a.x + row0 * a.ld, bytes, &full[stage]);
        if (HAS_RES) bulk_g2s(t0 + (AGG_TILE_BYTES / 4), a.res + row0 * a.ld, bytes, &full[stage]);
        if (HAS_DOT) bulk_g2s(t0 + (HAS_RES ? 2 : 1) * (AGG_TILE_BYTES / 4), a.dotx + row0 * a.ld, bytes, &full[stage]);
      }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ------------------------------------------------------------------ consumers: neighbour sums out of smem
    constexpr int RPW = 32 / LPR;                       // rows per warp step
    constexpr int res_off = HAS_RES ? (AGG_TILE_BYTES / 4) : 0;
    constexpr int dot_off = HAS_DOT ? (HAS_RES ? 2 : 1) * (AGG_TILE_BYTES / 4) : 0;
    const int cw = warp - 1;
    const int sub = lane / LPR, lc4 = (lane % LPR) * 4;
    const float one_eps = __fadd_rn(1.0f, a.eps ? __ldg(a.eps) : 0.0f);
    const int ld = a.ld;
    const bool one_pass = ld <= LPR * 4;               // every lane owns at most one float4 of its row
    double dot = 0.0;
    int stage = 0;
    uint32_t phase = 0;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
      mbar_wait(&full[stage], phase);
      const UnitDesc d = desc[stage];
      const float* tile = reinterpret_cast<const float*>(smem_raw + stage * stage_bytes);
      const uint32_t* nb = nbw + stage * AGG_NB_MAXN;
      float* out_tile = a.out + d.row0 * (long long)ld;
#pragma unroll 2
      for (int r = cw * RPW + sub; r < d.rows; r += AGG_CONSUMER_WARPS * RPW) {
        const int slot = (int)__umulhi((unsigned)r, d.magic);
        const int li = r - slot * d.n;
        const float* slot_tile = tile + slot * d.n * ld;
        const float* self_row = tile + r * ld;
        float* out_row = out_tile + r * ld;
        const uint32_t word = d.nb_local ? nb[li] : (NB_SLOW << 24);
        const uint32_t j0 = word & 0xFFu, j1 = (word >> 8) & 0xFFu, j2 = (word >> 16) & 0xFFu, j3 = word >> 24;
        if (j3 != NB_SLOW) {
          const float* p0 = j0 == NB_NONE ? zrow : slot_tile + j0 * ld;
          const float* p1 = j1 == NB_NONE ? zrow : slot_tile + j1 * ld;
          const float* p2 = j2 == NB_NONE ? zrow : slot_tile + j2 * ld;
          const float* p3 = j3 == NB_NONE ? zrow : slot_tile + j3 * ld;
          if (one_pass) {
            if (lc4 < ld)
              agg_cell<HAS_RES, HAS_DOT>(p0, p1, p2, p3, self_row, res_off, dot_off, one_eps, out_row, lc4, dot);
          } else {
            for (int col = lc4; col < ld; col += LPR * 4)
              agg_cell<HAS_RES, HAS_DOT>(p0, p1, p2, p3, self_row, res_off, dot_off, one_eps, out_row, col, dot);
          }
        } else {  // degree > 4 / large graph: walk the CSR in global memory (rows are still read from the tile)
          const int beg = __ldg(a.nbr_ptr + d.node0 + li), end = __ldg(a.nbr_ptr + d.node0 + li + 1);
          for (int col = lc4; col < ld; col += LPR * 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = beg; e < end; ++e) {
              const int lj = __ldg(a.nbr_idx + e) - d.node0;
              acc = add4(acc, lds4(slot_tile + lj * ld + col));
            }
            const float4 self = lds4(self_row + col);
            acc.x = __fadd_rn(acc.x, __fmul_rn(one_eps, self.x));
            acc.y = __fadd_rn(acc.y, __fmul_rn(one_eps, self.y));
            acc.z = __fadd_rn(acc.z, __fmul_rn(one_eps, self.z));
            acc.w = __fadd_rn(acc.w, __fmul_rn(one_eps, self.w));
            if (HAS_RES) {
              const float4 q = lds4(self_row + res_off + col);
              acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
            }
            if (HAS_DOT) {
              const float4 t = lds4(self_row + dot_off + col);
              dot += (double)(self.x * t.x + self.y * t.y + self.z * t.z + self.w * t.w);
            }
            stg4_stream(out_row + col, acc);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    if (HAS_DOT) {
      dot = warp_sum_d(dot);
      if (lane == 0) s_dot[cw] = dot;
    }
  }
  if (HAS_DOT) {
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < AGG_CONSUMER_WARPS; ++w) t += s_dot[w];
      atomicAdd(a.dot_out, t);
    }
  }
}

template <int LPR>
static int agg_launch_tma(const AggArgs& a, size_t smem, cudaStream_t st) {
  const int grid = sb_num_sms();
  const int mx = AGG_SMEM_LIMIT - 512;   // static shared (s_dot) counts against the 227 KB opt-in limit
#define AGG_GO(R_, D_)                                                                                          \
  do {                                                                                                          \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      SB_CUDA(cudaFuncSetAttribute(gin_agg_tma_kernel<LPR, R_, D_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   mx));                                                                        \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    gin_agg_tma_kernel<LPR, R_, D_><<<grid, AGG_THREADS, smem, st>>>(a);                                       \
  } while (0)
  if (a.res && a.dotx) AGG_GO(true, true);
  else if (a.res) AGG_GO(true, false);
  else if (a.dotx) AGG_GO(false, true);
  else AGG_GO(false, false);
#undef AGG_GO
  return SB_OK;
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
                          const int32_t* unit_desc, const uint32_t* nbr_pack, const int64_t* row_ptr,
                          const int32_t* nbr_ptr, const int32_t* nbr_idx, int64_t R, int32_t B, int32_t k,
                          int32_t masked, int32_t S, int32_t ld, int32_t tile_rows, int32_t force_generic,
                          void* stream) {
  SB_CHECK_ARG(R >= 0 && B >= 0 && S >= 1 && ld >= 1, "sb_gin_agg: bad sizes");
  SB_CHECK_ARG((dotx == nullptr) == (dot_out == nullptr), "sb_gin_agg: dotx and dot_out must be given together");
  SB_CHECK_ARG(x != out, "sb_gin_agg: x and out must not alias");
  if (R == 0 || B == 0) return SB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  AggArgs a;
  a.x = x; a.out = out; a.res = res; a.dotx = dotx; a.dot_out = dot_out; a.eps = eps;
  a.graph_ptr = graph_ptr; a.unit_ptr = unit_ptr; a.unit_desc = unit_desc; a.nbr_pack = nbr_pack; a.row_ptr = row_ptr; a.nbr_ptr = nbr_ptr; a.nbr_idx = nbr_idx;
  a.R = R; a.B = B; a.k = k; a.masked = masked; a.S = S; a.ld = ld; a.tile_rows = tile_rows; a.stages = 1;
  const bool tma_ok = !force_generic && unit_desc != nullptr && nbr_pack != nullptr && (ld % 4 == 0) &&
                      ld <= AGG_ZROW_FLOATS && tile_rows >= 1 &&
                      (int64_t)tile_rows * ld * 4 <= AGG_TILE_BYTES && ((uintptr_t)x % 16 == 0) &&
                      ((uintptr_t)out % 16 == 0) && (!res || (uintptr_t)res % 16 == 0) &&
                      (!dotx || (uintptr_t)dotx % 16 == 0);
  if (tma_ok) {
    const int T = 1 + (res ? 1 : 0) + (dotx ? 1 : 0);
    const size_t fixed = 2 * AGG_MAX_STAGES * sizeof(uint64_t) + AGG_MAX_STAGES * sizeof(UnitDesc) +
                         (size_t)AGG_MAX_STAGES * AGG_NB_MAXN * sizeof(uint32_t) + AGG_ZROW_FLOATS * sizeof(float);
    int stages = (int)((AGG_SMEM_LIMIT - 1024 - fixed) / ((size_t)T * AGG_TILE_BYTES));
    if (stages > AGG_MAX_STAGES) stages = AGG_MAX_STAGES;
    a.stages = stages;
    const size_t smem = (size_t)stages * T * AGG_TILE_BYTES + fixed;
    const int ld4 = ld / 4;
    const int lpr = ld4 > 16 ? 32 : ld4 > 8 ? 16 : ld4 > 4 ? 8 : 4;
    int rc;
    if (lpr == 32) rc = agg_launch_tma<32>(a, smem, st);
    else if (lpr == 16) rc = agg_launch_tma<16>(a, smem, st);
    else if (lpr == 8) rc = agg_launch_tma<8>(a, smem, st);
    else rc = agg_launch_tma<4>(a, smem, st);
    if (rc != SB_OK) return rc;
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

// Integer bookkeeping of the SignNet hot path (bit-exact rows a1/a2 of SURVEY.md §8): graph offsets from the sorted
// `batch` vector, the ragged slot-row layout, stable CSR/CSC of edge_index, and the re-layout of the ragged per-graph
// eigen-decomposition.  Replaces the torch_scatter / boolean-mask index math of
//   Alchemy/sign_net/transform.py:26-61 (to_dense_EVD, to_dense_list_EVD), sign_net.py:100-102 (mask build),
//   GraphPrediction/layers/deepsigns.py:66-78 (per-graph Python loop).
#include "common.cuh"
#include "../../include/signnet_b200.h"

#include <stdarg.h>

// ---------------------------------------------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";
void sb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* sb_last_error(void) { return g_err; }
extern "C" int sb_abi_version(void) { return SB_ABI_VERSION; }

int sb_num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = SB_NUM_SMS_FALLBACK;
  }
  return cached;
}
extern "C" int sb_device_sm_count(void) { return sb_num_sms(); }

// ------------------------------------------------------------------------------------------------------- graph_ptr
// batch is sorted non-decreasing with values in [0,B).  graph_ptr[b] = first node of graph b (empty graphs allowed).
__global__ void graph_ptr_kernel(const int64_t* __restrict__ batch, int64_t N, int B, int32_t* __restrict__ gp,
                                 int32_t* __restrict__ flags) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > N) return;
  int64_t cur = (i < N) ? batch[i] : (int64_t)B;
  int64_t prev = (i > 0) ? batch[i - 1] : -1;
  if (i < N && (cur < 0 || cur >= B)) {
    atomicOr(flags, SB_FLAG_BATCH_RANGE);
    return;
  }
  if (prev > cur) {
    atomicOr(flags, SB_FLAG_BATCH_UNSORTED);
    return;
  }
  if (prev < -1) prev = -1;
  for (int64_t b = prev + 1; b <= cur && b <= B; ++b) gp[b] = (int32_t)i;
}

extern "C" int sb_graph_ptr(const int64_t* batch, int64_t N, int32_t B, int32_t* graph_ptr, int32_t* flags,
                            void* stream) {
  SB_CHECK_ARG(N >= 0 && B >= 0 && N < (1ll << 31), "sb_graph_ptr: bad sizes N=%lld B=%d", (long long)N, B);
  cudaStream_t st = (cudaStream_t)stream;
  int threads = 256;
  int64_t blocks = sb_ceil_div(N + 1, threads);
  graph_ptr_kernel<<<(unsigned)blocks, threads, 0, st>>>(batch, N, B, graph_ptr, flags);
  SB_CHECK_LAUNCH("sb_graph_ptr");
  return SB_OK;
}

// ----------------------------------------------------------------------------------------------------- slot layout
// One block; sequential chunks with a running carry (B is the number of graphs: small).
//   k_b      = masked ? min(n_b, k) : k
//   row_ptr  = exclusive prefix of n_b * k_b      (slot-row offsets; row(b, j, i) = row_ptr[b] + j*n_b + i)
//   vec_ptr  = exclusive prefix of n_b * n_b      (offsets into the ragged eigen_vectors, transform.py:14)
//   unit_ptr = exclusive prefix of ceil(k_b / G_b), G_b = max(1, tile_rows / n_b)   (aggregate work units)
//   summary  = {R, n_max, k_b max, sum n_b^2, U, #graphs with n_b > tile_rows}
template <typename T>
__device__ T block_exclusive_scan_1024(T v, T* total, T* sh /*[33]*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    T s = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : (T)0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    sh[lane] = s;  // inclusive over warps
  }
  __syncthreads();
  T warp_off = (w > 0) ? sh[w - 1] : (T)0;
  *total = sh[(blockDim.x >> 5) - 1];
  T res = warp_off + x - v;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(1024) slot_layout_kernel(const int32_t* __restrict__ gp, int B, int k, int masked,
                                                           int tile_rows, int64_t* __restrict__ row_ptr,
                                                           int64_t* __restrict__ vec_ptr,
                                                           int32_t* __restrict__ unit_ptr,
                                                           int64_t* __restrict__ summary) {
  __shared__ long long sh[33];
  __shared__ int s_nmax, s_kmax, s_over;
  if (threadIdx.x == 0) s_nmax = 0, s_kmax = 0, s_over = 0;
  __syncthreads();
  long long carry_r = 0, carry_v = 0, carry_u = 0;
  for (int base = 0; base < B; base += blockDim.x) {
    int b = base + threadIdx.x;
    long long n = 0, kb = 0, units = 0;
    if (b < B) {
      n = gp[b + 1] - gp[b];
      kb = masked ? (n < k ? n : k) : k;
      if (n > 0) {
        long long g = tile_rows / n;
        if (g < 1) {
          g = 1;
          atomicAdd(&s_over, 1);
        }
        units = (kb + g - 1) / g;
      }
      atomicMax(&s_nmax, (int)n);
      if (n > 0) atomicMax(&s_kmax, (int)kb);
    }
    long long tot;
    long long er = block_exclusive_scan_1024<long long>(n * kb, &tot, sh);
    if (b < B) row_ptr[b] = carry_r + er;
    carry_r += tot;
    long long ev = block_exclusive_scan_1024<long long>(n * n, &tot, sh);
    if (b < B) vec_ptr[b] = carry_v + ev;
    carry_v += tot;
    long long eu = block_exclusive_scan_1024<long long>(units, &tot, sh);
    if (b < B) unit_ptr[b] = (int32_t)(carry_u + eu);
    carry_u += tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    row_ptr[B] = carry_r;
    vec_ptr[B] = carry_v;
    unit_ptr[B] = (int32_t)carry_u;
    summary[0] = carry_r;
    summary[1] = s_nmax;
    summary[2] = s_kmax;
    summary[3] = carry_v;
    summary[4] = carry_u;
    summary[5] = s_over;
  }
}

extern "C" int sb_slot_layout(const int32_t* graph_ptr, int32_t B, int32_t k, int32_t masked, int32_t tile_rows,
                              int64_t* row_ptr, int64_t* vec_ptr, int32_t* unit_ptr, int64_t* summary,
                              void* stream) {
  SB_CHECK_ARG(B >= 0 && k >= 1 && tile_rows >= 1, "sb_slot_layout: bad args B=%d k=%d tile_rows=%d", B, k,
               tile_rows);
  slot_layout_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(graph_ptr, B, k, masked, tile_rows, row_ptr, vec_ptr,
                                                           unit_ptr, summary);
  SB_CHECK_LAUNCH("sb_slot_layout");
  return SB_OK;
}

// Work units of the aggregate for another tile size (the slot layout itself does not depend on it).
__global__ void __launch_bounds__(1024) agg_units_kernel(const int32_t* __restrict__ gp, int B, int k, int masked,
                                                         int tile_rows, int32_t* __restrict__ unit_ptr) {
  __shared__ long long sh[33];
  long long carry = 0;
  for (int base = 0; base < B; base += blockDim.x) {
    int b = base + threadIdx.x;
    long long units = 0;
    if (b < B) {
      long long n = gp[b + 1] - gp[b];
      long long kb = masked ? (n < k ? n : k) : k;
      if (n > 0) {
        long long g = tile_rows / n;
        if (g < 1) g = 1;
        units = (kb + g - 1) / g;
      }
    }
    long long tot;
    long long e = block_exclusive_scan_1024<long long>(units, &tot, sh);
    if (b < B) unit_ptr[b] = (int32_t)(carry + e);
    carry += tot;
  }
  if (threadIdx.x == 0) unit_ptr[B] = (int32_t)carry;
}

extern "C" int sb_agg_units(const int32_t* graph_ptr, int32_t B, int32_t k, int32_t masked, int32_t tile_rows,
                            int32_t* unit_ptr, void* stream) {
  SB_CHECK_ARG(B >= 0 && k >= 1 && tile_rows >= 1, "sb_agg_units: bad args");
  agg_units_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(graph_ptr, B, k, masked, tile_rows, unit_ptr);
  SB_CHECK_LAUNCH("sb_agg_units");
  return SB_OK;
}

// Descriptor table of the aggregate's work units (one 48-byte record per (graph, slot-chunk) tile), so the producer
// warp of gin_agg_tma_kernel finds a tile with ONE coalesced load instead of a binary search + dependent pointer chase.
//   int32[12] = { row_rel lo, row_rel hi, n, rows, node0, magic = ceil(2^32/n), e0_in, ne_in, e0_out, ne_out, 0, 0 }
__global__ void agg_unit_desc_kernel(const int32_t* __restrict__ gp, const int64_t* __restrict__ row_ptr,
                                     const int32_t* __restrict__ unit_ptr, const int32_t* __restrict__ in_ptr,
                                     const int32_t* __restrict__ out_ptr, int B, int k, int masked, int tile_rows,
                                     int32_t* __restrict__ desc, long long cap) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int node0 = gp[b];
  const int n = gp[b + 1] - node0;
  if (n <= 0) return;
  const int kb = masked ? (n < k ? n : k) : k;
  int G = tile_rows / n;
  if (G < 1) G = 1;
  const int u0 = unit_ptr[b], u1 = unit_ptr[b + 1];
  const long long r0 = row_ptr[b];
  const int e0i = in_ptr[node0], nei = in_ptr[node0 + n] - e0i;
  const int e0o = out_ptr[node0], neo = out_ptr[node0 + n] - e0o;
  const unsigned magic = (unsigned)((0x100000000ull + (unsigned)n - 1) / (unsigned)n);
  for (int u = u0; u < u1 && u < cap; ++u) {
    const int j0 = (u - u0) * G;
    const int ns = (kb - j0 < G) ? (kb - j0) : G;
    const long long rr = r0 + (long long)j0 * n;
    int32_t* d = desc + (long long)u * 12;
    d[0] = (int32_t)(rr & 0xffffffffll);
    d[1] = (int32_t)(rr >> 32);
    d[2] = n;
    d[3] = ns * n;
    d[4] = node0;
    d[5] = (int32_t)magic;
    d[6] = e0i; d[7] = nei; d[8] = e0o; d[9] = neo;
    d[10] = 0; d[11] = 0;
  }
}

extern "C" int sb_agg_unit_desc(const int32_t* graph_ptr, const int64_t* row_ptr, const int32_t* unit_ptr,
                                const int32_t* in_ptr, const int32_t* out_ptr, int32_t B, int32_t k, int32_t masked,
                                int32_t tile_rows, int32_t* unit_desc, int64_t cap_units, void* stream) {
  SB_CHECK_ARG(B >= 0 && k >= 1 && tile_rows >= 1 && cap_units >= 0, "sb_agg_unit_desc: bad args");
  if (B == 0) return SB_OK;
  agg_unit_desc_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(graph_ptr, row_ptr, unit_ptr, in_ptr,
                                                                          out_ptr, B, k, masked, tile_rows, unit_desc,
                                                                          cap_units);
  SB_CHECK_LAUNCH("sb_agg_unit_desc");
  return SB_OK;
}

// Packed neighbour word per node for the TMA aggregate: up to four LOCAL neighbour ids (node id - first node of the
// graph) in CSR order, one per byte, 0xFF = empty.  Byte 3 = 0xFE marks a node the fast path cannot express (degree
// > 4 or a local id > 253): the aggregate then walks the CSR for that node.
__global__ void pack_neighbours_kernel(const int64_t* __restrict__ batch, const int32_t* __restrict__ gp,
                                       const int32_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr_idx,
                                       long long N, uint32_t* __restrict__ pack) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int node0 = gp[batch[i]];
  const int beg = nbr_ptr[i], end = nbr_ptr[i + 1];
  uint32_t w = 0xFFFFFFFFu;
  bool slow = end - beg > 4;
  if (!slow) {
    for (int e = beg; e < end; ++e) {
      const int lj = nbr_idx[e] - node0;
      if (lj < 0 || lj > 253) { slow = true; break; }
      const int q = e - beg;
      w = (w & ~(0xFFu << (8 * q))) | ((uint32_t)lj << (8 * q));
    }
  }
  pack[i] = slow ? 0xFEFFFFFFu : w;
}

extern "C" int sb_pack_neighbours(const int64_t* batch, const int32_t* graph_ptr, const int32_t* nbr_ptr,
                                  const int32_t* nbr_idx, int64_t N, uint32_t* nbr_pack, void* stream) {
  SB_CHECK_ARG(N >= 0, "sb_pack_neighbours: bad N");
  if (N == 0) return SB_OK;
  pack_neighbours_kernel<<<(unsigned)sb_ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(batch, graph_ptr, nbr_ptr,
                                                                                          nbr_idx, N, nbr_pack);
  SB_CHECK_LAUNCH("sb_pack_neighbours");
  return SB_OK;
}

// ------------------------------------------------------------------------------------------------------ device scan
// exclusive scan of int32 counts (n elements, in place), n <= 4096*4096.
#define SCAN_ITEMS 4
#define SCAN_BLOCK 1024
#define SCAN_TILE (SCAN_ITEMS * SCAN_BLOCK)
__global__ void __launch_bounds__(SCAN_BLOCK) scan_tiles_kernel(int32_t* data, int64_t n, int32_t* partials) {
  __shared__ int sh[33];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? data[base + i] : 0;
    s += v[i];
  }
  int tot;
  int e = block_exclusive_scan_1024<int>(s, &tot, sh);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) data[base + i] = e;
    e += v[i];
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_partials_kernel(int32_t* partials, int nparts) {
  __shared__ int sh[33];
  int carry = 0;
  for (int base = 0; base < nparts; base += SCAN_BLOCK) {
    int i = base + threadIdx.x;
    int v = (i < nparts) ? partials[i] : 0;
    int tot;
    int e = block_exclusive_scan_1024<int>(v, &tot, sh);
    if (i < nparts) partials[i] = carry + e;
    carry += tot;
  }
}
__global__ void scan_add_kernel(int32_t* data, int64_t n, const int32_t* partials) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) data[i] += partials[i / SCAN_TILE];
}
static int exclusive_scan_i32(int32_t* data, int64_t n, int32_t* partials, cudaStream_t st) {
  int64_t tiles = sb_ceil_div(n, SCAN_TILE);
  if (tiles == 0) return SB_OK;
  scan_tiles_kernel<<<(unsigned)tiles, SCAN_BLOCK, 0, st>>>(data, n, partials);
  if (tiles > 1) {
    scan_partials_kernel<<<1, SCAN_BLOCK, 0, st>>>(partials, (int)tiles);
    scan_add_kernel<<<(unsigned)sb_ceil_div(n, 256), 256, 0, st>>>(data, n, partials);
  }
  return SB_OK;
}

// ------------------------------------------------------------------------------------------------- stable CSR / CSC
__global__ void csr_count_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                 int64_t N, const int64_t* __restrict__ batch, int32_t* __restrict__ in_cnt,
                                 int32_t* __restrict__ out_cnt, int32_t* __restrict__ flags) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = src[e], d = dst[e];
  if (s < 0 || s >= N || d < 0 || d >= N) {
    atomicOr(flags, SB_FLAG_EDGE_RANGE);
    return;
  }
  if (batch != nullptr && batch[s] != batch[d]) atomicOr(flags, SB_FLAG_EDGE_CROSS_GRAPH);
  atomicAdd(&in_cnt[d], 1);
  atomicAdd(&out_cnt[s], 1);
}
__global__ void csr_fill_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E,
                                int64_t N, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ out_ptr,
                                int32_t* __restrict__ in_cur, int32_t* __restrict__ out_cur,
                                int32_t* __restrict__ in_eid, int32_t* __restrict__ out_eid) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = src[e], d = dst[e];
  if (s < 0 || s >= N || d < 0 || d >= N) return;
  in_eid[in_ptr[d] + atomicAdd(&in_cur[d], 1)] = (int32_t)e;
  out_eid[out_ptr[s] + atomicAdd(&out_cur[s], 1)] = (int32_t)e;
}
// Restore edge-id order inside every row (the atomics above hand out slots in arbitrary order) so neighbour sums
// are accumulated in exactly the order torch's CPU index_add_ / scatter_add uses, then resolve the other endpoint.
__global__ void csr_sort_rows_kernel(const int64_t* __restrict__ other, int64_t N, const int32_t* __restrict__ ptr,
                                     int32_t* __restrict__ eid, int32_t* __restrict__ nbr) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= N) return;
  int lo = ptr[v], hi = ptr[v + 1];
  for (int i = lo + 1; i < hi; ++i) {
    int key = eid[i], j = i - 1;
    while (j >= lo && eid[j] > key) {
      eid[j + 1] = eid[j];
      --j;
    }
    eid[j + 1] = key;
  }
  for (int i = lo; i < hi; ++i) nbr[i] = (int32_t)other[eid[i]];
}

extern "C" int sb_build_csr(const int64_t* edge_index, int64_t E, int64_t N, const int64_t* batch, int32_t* in_ptr,
                            int32_t* in_src, int32_t* in_eid, int32_t* out_ptr, int32_t* out_dst, int32_t* out_eid,
                            int32_t* workspace, int64_t workspace_ints, int32_t* flags, void* stream) {
  SB_CHECK_ARG(E >= 0 && N >= 0 && N < (1ll << 31) - 1 && E < (1ll << 31), "sb_build_csr: bad sizes");
  int64_t tiles = sb_ceil_div(N + 1, SCAN_TILE);
  SB_CHECK_ARG(tiles <= SCAN_TILE, "sb_build_csr: N too large for the two-level scan");
  int64_t need = 2 * (N + 1) + 2 * tiles;
  SB_CHECK_ARG(workspace_ints >= need, "sb_build_csr: workspace too small (%lld < %lld ints)",
               (long long)workspace_ints, (long long)need);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + E;
  int32_t* in_cur = workspace;
  int32_t* out_cur = workspace + (N + 1);
  int32_t* part_a = workspace + 2 * (N + 1);
  int32_t* part_b = part_a + tiles;
  SB_CUDA(cudaMemsetAsync(in_ptr, 0, sizeof(int32_t) * (N + 1), st));
  SB_CUDA(cudaMemsetAsync(out_ptr, 0, sizeof(int32_t) * (N + 1), st));
  SB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(int32_t) * 2 * (N + 1), st));
  if (E > 0) {
    csr_count_kernel<<<(unsigned)sb_ceil_div(E, 256), 256, 0, st>>>(src, dst, E, N, batch, in_ptr, out_ptr, flags);
    SB_CHECK_LAUNCH("csr_count");
  }
  exclusive_scan_i32(in_ptr, N + 1, part_a, st);
  exclusive_scan_i32(out_ptr, N + 1, part_b, st);
  SB_CHECK_LAUNCH("csr_scan");
  if (E > 0) {
    csr_fill_kernel<<<(unsigned)sb_ceil_div(E, 256), 256, 0, st>>>(src, dst, E, N, in_ptr, out_ptr, in_cur, out_cur,
                                                                   in_eid, out_eid);
    SB_CHECK_LAUNCH("csr_fill");
  }
  if (N > 0) {
    csr_sort_rows_kernel<<<(unsigned)sb_ceil_div(N, 128), 128, 0, st>>>(src, N, in_ptr, in_eid, in_src);
    csr_sort_rows_kernel<<<(unsigned)sb_ceil_div(N, 128), 128, 0, st>>>(dst, N, out_ptr, out_eid, out_dst);
    SB_CHECK_LAUNCH("csr_sort_rows");
  }
  return SB_OK;
}

// -------------------------------------------------------------------------------- ragged EVD -> phi input / dense list
// x0[s, row_ptr[b] + j*n_b + i] = (+1,-1)[s] * V_b[i, j]   for j < k_b   (V_b row-major [node, eig]).
__global__ void phi_input_ragged_kernel(const float* __restrict__ evec, const int64_t* __restrict__ batch,
                                        const int32_t* __restrict__ gp, const int64_t* __restrict__ row_ptr,
                                        const int64_t* __restrict__ vec_ptr, int64_t N, int k, int masked, int64_t R,
                                        float* __restrict__ x0) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * (int64_t)k) return;
  int64_t node = t / k;
  int j = (int)(t % k);
  int b = (int)batch[node];
  int n = gp[b + 1] - gp[b];
  int kb = masked ? (n < k ? n : k) : k;
  if (j >= kb) return;
  int li = (int)(node - gp[b]);
  float v = (j < n) ? evec[vec_ptr[b] + (int64_t)li * n + j] : 0.f;  // unmasked layouts pad missing columns with 0
  int64_t r = row_ptr[b] + (int64_t)j * n + li;
  x0[r] = v;
  x0[R + r] = -v;
}
// Same from the dense-list tensor eigvecs [N, kd] (kd >= k columns available).
__global__ void phi_input_dense_kernel(const float* __restrict__ eigvecs, int64_t ld, const int64_t* __restrict__ batch,
                                       const int32_t* __restrict__ gp, const int64_t* __restrict__ row_ptr,
                                       int64_t N, int k, int masked, int64_t R, float* __restrict__ x0) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * (int64_t)k) return;
  int64_t node = t / k;
  int j = (int)(t % k);
  int b = (int)batch[node];
  int n = gp[b + 1] - gp[b];
  int kb = masked ? (n < k ? n : k) : k;
  if (j >= kb) return;
  int li = (int)(node - gp[b]);
  float v = eigvecs[node * ld + j];
  int64_t r = row_ptr[b] + (int64_t)j * n + li;
  x0[r] = v;
  x0[R + r] = -v;
}

extern "C" int sb_phi_input_ragged(const float* eigen_vectors, const int64_t* batch, const int32_t* graph_ptr,
                                   const int64_t* row_ptr, const int64_t* vec_ptr, int64_t N, int32_t k,
                                   int32_t masked, int64_t R, float* x0, void* stream) {
  if (N == 0) return SB_OK;
  int64_t tot = N * (int64_t)k;
  phi_input_ragged_kernel<<<(unsigned)sb_ceil_div(tot, 256), 256, 0, (cudaStream_t)stream>>>(
      eigen_vectors, batch, graph_ptr, row_ptr, vec_ptr, N, k, masked, R, x0);
  SB_CHECK_LAUNCH("sb_phi_input_ragged");
  return SB_OK;
}
extern "C" int sb_phi_input_dense(const float* eigvecs, int64_t ld, const int64_t* batch, const int32_t* graph_ptr,
                                  const int64_t* row_ptr, int64_t N, int32_t k, int32_t masked, int64_t R, float* x0,
                                  void* stream) {
  if (N == 0) return SB_OK;
  int64_t tot = N * (int64_t)k;
  phi_input_dense_kernel<<<(unsigned)sb_ceil_div(tot, 256), 256, 0, (cudaStream_t)stream>>>(
      eigvecs, ld, batch, graph_ptr, row_ptr, N, k, masked, R, x0);
  SB_CHECK_LAUNCH("sb_phi_input_dense");
  return SB_OK;
}

// Per-slot-row scalar broadcast of a per-node vector laid out per graph, e.g. the eigenvalue feature of the
// eigen_encoder (sign_net.py:107-108): out[row(b,j,i)] = eigen_values[gp[b] + j]   (eigenvalue j of graph b).
__global__ void slot_eigval_kernel(const float* __restrict__ eval, const int64_t* __restrict__ batch,
                                   const int32_t* __restrict__ gp, const int64_t* __restrict__ row_ptr, int64_t N,
                                   int k, int masked, float* __restrict__ out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * (int64_t)k) return;
  int64_t node = t / k;
  int j = (int)(t % k);
  int b = (int)batch[node];
  int n = gp[b + 1] - gp[b];
  int kb = masked ? (n < k ? n : k) : k;
  if (j >= kb) return;
  int li = (int)(node - gp[b]);
  out[row_ptr[b] + (int64_t)j * n + li] = (j < n) ? eval[gp[b] + j] : 0.f;
}
extern "C" int sb_slot_eigval(const float* eigen_values, const int64_t* batch, const int32_t* graph_ptr,
                              const int64_t* row_ptr, int64_t N, int32_t k, int32_t masked, float* out,
                              void* stream) {
  if (N == 0) return SB_OK;
  int64_t tot = N * (int64_t)k;
  slot_eigval_kernel<<<(unsigned)sb_ceil_div(tot, 256), 256, 0, (cudaStream_t)stream>>>(
      eigen_values, batch, graph_ptr, row_ptr, N, k, masked, out);
  SB_CHECK_LAUNCH("sb_slot_eigval");
  return SB_OK;
}

// to_dense_list_EVD (transform.py:52-61) straight from the ragged inputs, no [B,Nmax,Nmax] detour:
//   eigS[i, j] = lambda_{b(i), j},  eigV[i, j] = V_b[local(i), j]   for j < n_b, else 0;  mask[i, j] = j < n_b.
__global__ void dense_list_evd_kernel(const float* __restrict__ eval, const float* __restrict__ evec,
                                      const int64_t* __restrict__ batch, const int32_t* __restrict__ gp,
                                      const int64_t* __restrict__ vec_ptr, int64_t N, int nmax,
                                      float* __restrict__ eigS, float* __restrict__ eigV,
                                      uint8_t* __restrict__ mask) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * (int64_t)nmax) return;
  int64_t node = t / nmax;
  int j = (int)(t % nmax);
  int b = (int)batch[node];
  int n = gp[b + 1] - gp[b];
  int li = (int)(node - gp[b]);
  bool ok = j < n;
  if (eigS) eigS[t] = ok ? eval[gp[b] + j] : 0.f;
  if (eigV) eigV[t] = ok ? evec[vec_ptr[b] + (int64_t)li * n + j] : 0.f;
  if (mask) mask[t] = ok ? 1 : 0;
}
extern "C" int sb_dense_list_evd(const float* eigen_values, const float* eigen_vectors, const int64_t* batch,
                                 const int32_t* graph_ptr, const int64_t* vec_ptr, int64_t N, int32_t nmax,
                                 float* eigS, float* eigV, uint8_t* mask, void* stream) {
  if (N == 0 || nmax == 0) return SB_OK;
  int64_t tot = N * (int64_t)nmax;
  dense_list_evd_kernel<<<(unsigned)sb_ceil_div(tot, 256), 256, 0, (cudaStream_t)stream>>>(
      eigen_values, eigen_vectors, batch, graph_ptr, vec_ptr, N, nmax, eigS, eigV, mask);
  SB_CHECK_LAUNCH("sb_dense_list_evd");
  return SB_OK;
}

// ------------------------------------------------------------------------------ slot rows <-> dense [N, k, C] tensors
// dense[node, j, c] = sum_s rows[s, row(b,j,i), c]  (0 for slots j >= k_b): the reference's padded [N,k,d] view, e.g.
// the return value of GNN3d (sign_net.py:44) or phi(x)+phi(-x) (sign_net.py:113 / deepsigns.py:73).
__global__ void rows_to_dense_kernel(const float* __restrict__ rows, long long ld, long long R, int S,
                                     const int64_t* __restrict__ batch, const int32_t* __restrict__ gp,
                                     const int64_t* __restrict__ row_ptr, long long N, int k, int masked, int C,
                                     float* __restrict__ dense) {
  const long long total = N * (long long)k * C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const long long nj = t / C;
    const int j = (int)(nj % k);
    const long long node = nj / k;
    const int b = (int)batch[node];
    const int n = gp[b + 1] - gp[b];
    const int kb = masked ? (n < k ? n : k) : k;
    float v = 0.f;
    if (j < kb) {
      const long long r = row_ptr[b] + (long long)j * n + (node - gp[b]);
      for (int s = 0; s < S; ++s) v += __ldg(rows + ((long long)s * R + r) * ld + c);
    }
    dense[t] = v;
  }
}
extern "C" int sb_rows_to_dense(const float* rows, int64_t ld, int64_t R, int32_t S, const int64_t* batch,
                                const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N, int32_t k,
                                int32_t masked, int32_t C, float* dense, void* stream) {
  const long long total = N * (long long)k * C;
  if (total == 0) return SB_OK;
  long long blocks = sb_ceil_div(total, 256);
  const long long cap = (long long)sb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  rows_to_dense_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rows, ld, R, S, batch, graph_ptr, row_ptr,
                                                                          N, k, masked, C, dense);
  SB_CHECK_LAUNCH("sb_rows_to_dense");
  return SB_OK;
}

// rows[s, row(b,j,i), c] = sign_s * dense[node, j, c]  for c < C, 0 for the padding columns C..ld-1
// (signs: S=1 -> {+1}; S=2 -> {+1, -1} when negate_second, else {+1, +1}: the backward of the sum over s).
__global__ void dense_to_rows_kernel(const float* __restrict__ dense, long long R, int S, int negate_second,
                                     const int64_t* __restrict__ batch, const int32_t* __restrict__ gp,
                                     const int64_t* __restrict__ row_ptr, long long N, int k, int masked, int C,
                                     long long ld, float* __restrict__ rows) {
  const long long total = N * (long long)k * ld;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % ld);
    const long long nj = t / ld;
    const int j = (int)(nj % k);
    const long long node = nj / k;
    const int b = (int)batch[node];
    const int n = gp[b + 1] - gp[b];
    const int kb = masked ? (n < k ? n : k) : k;
    if (j >= kb) continue;
    const long long r = row_ptr[b] + (long long)j * n + (node - gp[b]);
    const float v = (c < C) ? __ldg(dense + (node * k + j) * C + c) : 0.f;
    rows[r * ld + c] = v;
    if (S == 2) rows[(R + r) * ld + c] = negate_second ? -v : v;
  }
}
extern "C" int sb_dense_to_rows(const float* dense, int64_t R, int32_t S, int32_t negate_second,
                                const int64_t* batch, const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N,
                                int32_t k, int32_t masked, int32_t C, int64_t ld, float* rows, void* stream) {
  SB_CHECK_ARG(S == 1 || S == 2, "sb_dense_to_rows: S must be 1 or 2");
  const long long total = N * (long long)k * ld;
  if (total == 0) return SB_OK;
  long long blocks = sb_ceil_div(total, 256);
  const long long cap = (long long)sb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  dense_to_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dense, R, S, negate_second, batch,
                                                                          graph_ptr, row_ptr, N, k, masked, C, ld,
                                                                          rows);
  SB_CHECK_LAUNCH("sb_dense_to_rows");
  return SB_OK;
}

// K5/K6 — the downstream GINE predictor's sparse pieces on [N, ld] node rows:
//   GINE aggregate  out_i = (1+eps) x_i + sum_{(j->i)} relu(x_j + e_ji)       (pyg_gnn_wrapper.py:19-28, PyG GINEConv)
//   add / mean pool out_b = sum_{i in graph b} x_i                             (model.py:58-61, torch_scatter.scatter)
//   DiscreteEncoder out_m = sum_f Embedding_f[idx[m, f]]                       (elements.py:21-37)
// One warp per destination row, float4 lanes over the feature dim, neighbours visited in stable CSR (= edge id) order so
// the fp32 sums are bit-identical to the CPU reference's index_add_; no atomics on the forward, deterministic.
// x (N <= ~25k rows x 512 B) is L2-resident on B200, so neighbour rows are read straight through L1/L2.
#include "common.cuh"
#include "../../include/signnet_b200.h"

__global__ void __launch_bounds__(256) gine_agg_fwd_kernel(const float* __restrict__ x, const float* __restrict__ e,
                                                           const float* __restrict__ eps,
                                                           const int32_t* __restrict__ in_ptr,
                                                           const int32_t* __restrict__ in_src,
                                                           const int32_t* __restrict__ in_eid, long long N, int ld,
                                                           float* __restrict__ out) {
  const long long node = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= N) return;
  const float one_eps = __fadd_rn(1.0f, eps ? __ldg(eps) : 0.0f);
  const int beg = __ldg(in_ptr + node), end = __ldg(in_ptr + node + 1);
  for (int c4 = lane; c4 < (ld >> 2); c4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = beg; p < end; ++p) {
      const float4 xv = ldg4(x + (long long)__ldg(in_src + p) * ld + c4 * 4);
      const float4 ev = ldg4(e + (long long)__ldg(in_eid + p) * ld + c4 * 4);
      acc.x = __fadd_rn(acc.x, fmaxf(__fadd_rn(xv.x, ev.x), 0.f));
      acc.y = __fadd_rn(acc.y, fmaxf(__fadd_rn(xv.y, ev.y), 0.f));
      acc.z = __fadd_rn(acc.z, fmaxf(__fadd_rn(xv.z, ev.z), 0.f));
      acc.w = __fadd_rn(acc.w, fmaxf(__fadd_rn(xv.w, ev.w), 0.f));
    }
    const float4 s = ldg4(x + node * ld + c4 * 4);
    acc.x = __fadd_rn(acc.x, __fmul_rn(one_eps, s.x));
    acc.y = __fadd_rn(acc.y, __fmul_rn(one_eps, s.y));
    acc.z = __fadd_rn(acc.z, __fmul_rn(one_eps, s.z));
    acc.w = __fadd_rn(acc.w, __fmul_rn(one_eps, s.w));
    *reinterpret_cast<float4*>(out + node * ld + c4 * 4) = acc;
  }
}

extern "C" int sb_gine_agg_fwd(const float* x, const float* e, const float* eps, const int32_t* in_ptr,
                               const int32_t* in_src, const int32_t* in_eid, int64_t N, int32_t ld, float* out,
                               void* stream) {
  SB_CHECK_ARG(ld % 4 == 0 && ld > 0, "sb_gine_agg_fwd: ld must be a positive multiple of 4");
  if (N == 0) return SB_OK;
  gine_agg_fwd_kernel<<<(unsigned)sb_ceil_div(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, e, eps, in_ptr, in_src,
                                                                                          in_eid, N, ld, out);
  SB_CHECK_LAUNCH("sb_gine_agg_fwd");
  return SB_OK;
}

// backward, node part:  dx_i = (1+eps) dA_i + sum_{(i->t)} dA_t * [x_i + e_it > 0] ;  deps += sum dA * x
__global__ void __launch_bounds__(256) gine_agg_bwd_node_kernel(const float* __restrict__ dA,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ e,
                                                                const float* __restrict__ eps,
                                                                const int32_t* __restrict__ out_ptr,
                                                                const int32_t* __restrict__ out_dst,
                                                                const int32_t* __restrict__ out_eid, long long N,
                                                                int ld, float* __restrict__ dx,
                                                                double* __restrict__ deps) {
  const long long node = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  double dot = 0.0;
  if (node < N) {
    const float one_eps = 1.0f + (eps ? __ldg(eps) : 0.0f);
    const int beg = __ldg(out_ptr + node), end = __ldg(out_ptr + node + 1);
    for (int c4 = lane; c4 < (ld >> 2); c4 += 32) {
      const float4 xs = ldg4(x + node * ld + c4 * 4);
      const float4 gs = ldg4(dA + node * ld + c4 * 4);
      float4 acc = make_float4(one_eps * gs.x, one_eps * gs.y, one_eps * gs.z, one_eps * gs.w);
      for (int p = beg; p < end; ++p) {
        const float4 g = ldg4(dA + (long long)__ldg(out_dst + p) * ld + c4 * 4);
        const float4 ev = ldg4(e + (long long)__ldg(out_eid + p) * ld + c4 * 4);
        if (__fadd_rn(xs.x, ev.x) > 0.f) acc.x += g.x;
        if (__fadd_rn(xs.y, ev.y) > 0.f) acc.y += g.y;
        if (__fadd_rn(xs.z, ev.z) > 0.f) acc.z += g.z;
        if (__fadd_rn(xs.w, ev.w) > 0.f) acc.w += g.w;
      }
      *reinterpret_cast<float4*>(dx + node * ld + c4 * 4) = acc;
      dot += (double)(gs.x * xs.x + gs.y * xs.y + gs.z * xs.z + gs.w * xs.w);
    }
  }
  if (deps) {
    dot = warp_sum_d(dot);
    if (lane == 0 && dot != 0.0) atomicAdd(deps, dot);
  }
}
// backward, edge part:  de_k = dA_{dst_k} * [x_{src_k} + e_k > 0]   (one warp per edge)
__global__ void __launch_bounds__(256) gine_agg_bwd_edge_kernel(const float* __restrict__ dA,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ e,
                                                                const int64_t* __restrict__ src,
                                                                const int64_t* __restrict__ dst, long long E, int ld,
                                                                float* __restrict__ de) {
  const long long k = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= E) return;
  const long long s = src[k], t = dst[k];
  for (int c4 = lane; c4 < (ld >> 2); c4 += 32) {
    const float4 xs = ldg4(x + s * ld + c4 * 4);
    const float4 ev = ldg4(e + k * ld + c4 * 4);
    const float4 g = ldg4(dA + t * ld + c4 * 4);
    float4 o;
    o.x = (__fadd_rn(xs.x, ev.x) > 0.f) ? g.x : 0.f;
    o.y = (__fadd_rn(xs.y, ev.y) > 0.f) ? g.y : 0.f;
    o.z = (__fadd_rn(xs.z, ev.z) > 0.f) ? g.z : 0.f;
    o.w = (__fadd_rn(xs.w, ev.w) > 0.f) ? g.w : 0.f;
    *reinterpret_cast<float4*>(de + k * ld + c4 * 4) = o;
  }
}

extern "C" int sb_gine_agg_bwd(const float* dA, const float* x, const float* e, const float* eps,
                               const int64_t* edge_index, const int32_t* out_ptr, const int32_t* out_dst,
                               const int32_t* out_eid, int64_t N, int64_t E, int32_t ld, float* dx, float* de,
                               double* deps, void* stream) {
  SB_CHECK_ARG(ld % 4 == 0 && ld > 0, "sb_gine_agg_bwd: ld must be a positive multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  if (N > 0 && dx) {
    gine_agg_bwd_node_kernel<<<(unsigned)sb_ceil_div(N * 32, 256), 256, 0, st>>>(dA, x, e, eps, out_ptr, out_dst,
                                                                               out_eid, N, ld, dx, deps);
    SB_CHECK_LAUNCH("sb_gine_agg_bwd(node)");
  }
  if (E > 0 && de) {
    gine_agg_bwd_edge_kernel<<<(unsigned)sb_ceil_div(E * 32, 256), 256, 0, st>>>(dA, x, e, edge_index, edge_index + E,
                                                                               E, ld, de);
    SB_CHECK_LAUNCH("sb_gine_agg_bwd(edge)");
  }
  return SB_OK;
}

// ----------------------------------------------------------------------------------------------------- graph pooling
// out[b, :] = sum_{i in graph b} x[i, :]  (optionally / max(n_b, 1)); node order = CPU index_add_ order.
__global__ void __launch_bounds__(256) segment_pool_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                               const int32_t* __restrict__ gp, int B, int C, int mean,
                                                               float* __restrict__ out, long long ldo) {
  const long long b = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int beg = gp[b], end = gp[b + 1];
  for (int c = lane; c < (int)ldo; c += 32) {
    float acc = 0.f;
    if (c < C)
      for (int i = beg; i < end; ++i) acc = __fadd_rn(acc, __ldg(x + (long long)i * ldx + c));
    if (mean && c < C) acc = acc / (float)((end - beg) > 1 ? (end - beg) : 1);
    out[b * ldo + c] = acc;
  }
}
extern "C" int sb_segment_pool_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int32_t B, int32_t C,
                                   int32_t mean, float* out, int64_t ldo, void* stream) {
  if (B == 0) return SB_OK;
  segment_pool_fwd_kernel<<<(unsigned)sb_ceil_div((long long)B * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, graph_ptr, B, C, mean, out, ldo);
  SB_CHECK_LAUNCH("sb_segment_pool_fwd");
  return SB_OK;
}
__global__ void segment_pool_bwd_kernel(const float* __restrict__ gout, long long ldo, const int64_t* __restrict__ batch,
                                        const int32_t* __restrict__ gp, long long N, int C, int mean,
                                        float* __restrict__ gx, long long ldx) {
  const long long total = N * ldx;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / ldx;
    const int c = (int)(t - i * ldx);
    float v = 0.f;
    if (c < C) {
      const int b = (int)batch[i];
      v = __ldg(gout + (long long)b * ldo + c);
      if (mean) {
        const int n = gp[b + 1] - gp[b];
        v = v / (float)(n > 1 ? n : 1);
      }
    }
    gx[t] = v;
  }
}
extern "C" int sb_segment_pool_bwd(const float* gout, int64_t ldo, const int64_t* batch, const int32_t* graph_ptr,
                                   int64_t N, int32_t C, int32_t mean, float* gx, int64_t ldx, void* stream) {
  if (N == 0) return SB_OK;
  long long blocks = sb_ceil_div(N * ldx, 256);
  const long long cap = (long long)sb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  segment_pool_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(gout, ldo, batch, graph_ptr, N, C, mean,
                                                                             gx, ldx);
  SB_CHECK_LAUNCH("sb_segment_pool_bwd");
  return SB_OK;
}

// ------------------------------------------------------------------------------------------ embedding sum (discrete)
// out[m, :] (+)= table[idx[m*stride], :]   (one feature column per call; DiscreteEncoder sums over columns)
__global__ void embedding_fwd_kernel(const int64_t* __restrict__ idx, long long stride, const float* __restrict__ table,
                                     int V, int C, long long M, float* __restrict__ out, long long ldo, int accumulate,
                                     int32_t* __restrict__ flags) {
  const long long total = M * ldo;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / ldo;
    const int c = (int)(t - m * ldo);
    const long long v = idx[m * stride];
    float val = 0.f;
    if (v < 0 || v >= V) {
      // nn.Embedding raises for such an index (CPU: IndexError; CUDA: device-side assert).  Same here: the flag is
      // set for callers that read it, and the kernel traps - the next CUDA call of the process reports the failure
      // instead of training silently on zero embeddings.
      if (c == 0) {
        if (flags) atomicOr(flags, 1);
        printf("libsignnet_b200: DiscreteEncoder index %lld out of range [0, %d) at row %lld\n", v, V, m);
      }
      __trap();
    } else if (c < C) {
      val = __ldg(table + v * C + c);
    }
    out[t] = accumulate ? out[t] + val : val;
  }
}
extern "C" int sb_embedding_fwd(const int64_t* idx, int64_t stride, const float* table, int32_t V, int32_t C, int64_t M,
                                float* out, int64_t ldo, int32_t accumulate, int32_t* flags, void* stream) {
  if (M == 0) return SB_OK;
  long long blocks = sb_ceil_div(M * ldo, 256);
  const long long cap = (long long)sb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  embedding_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(idx, stride, table, V, C, M, out, ldo,
                                                                          accumulate, flags);
  SB_CHECK_LAUNCH("sb_embedding_fwd");
  return SB_OK;
}
// dtable[v, c] = sum_{m: idx[m] = v} g[m, c].  Deterministic: CTA p owns a contiguous chunk of rows, thread c owns
// column c and walks the chunk in order into a private [V] column held in shared memory; the per-CTA partial tables
// are then added in CTA order by the second kernel.
__global__ void embedding_bwd_partial_kernel(const int64_t* __restrict__ idx, long long stride,
                                             const float* __restrict__ g, long long ldg, int V, int C, long long M,
                                             float* __restrict__ partial) {
  extern __shared__ float tbl[];  // [V][blockDim.x]
  const int c = threadIdx.x;
  for (int v = 0; v < V; ++v) tbl[v * blockDim.x + c] = 0.f;
  const long long chunk = (M + gridDim.x - 1) / gridDim.x;
  const long long beg = blockIdx.x * chunk, end = (beg + chunk < M) ? beg + chunk : M;
  if (c < C)
    for (long long m = beg; m < end; ++m) {
      const long long v = idx[m * stride];
      if (v >= 0 && v < V) tbl[v * blockDim.x + c] += __ldg(g + m * ldg + c);
    }
  if (c < C)
    for (int v = 0; v < V; ++v) partial[((long long)blockIdx.x * V + v) * C + c] = tbl[v * blockDim.x + c];
}
__global__ void embedding_bwd_reduce_kernel(const float* __restrict__ partial, int nparts, int V, int C,
                                            float* __restrict__ dtable) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V * C) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(long long)p * V * C + t];
  dtable[t] = s;
}
// Large vocabularies (table does not fit shared memory): fp32 atomics into a zeroed table.  Categorical features use
// few distinct values (28 atom / 4 bond types inside DiscreteEncoder's 500-row tables, elements.py:22-25), so rows
// v < EMB_HOT are privatised per CTA in shared memory: thread c owns column c and walks the CTA's rows in order, so the
// private table needs no atomics at all (the first version used shared-memory atomics: 2 cycles per lane, 86 us for
// 50 k edges x 128 columns); one global atomic per touched element and CTA at the end.  Colder rows go straight to global
// memory.
#define EMB_HOT 64
__global__ void __launch_bounds__(128) embedding_bwd_hot_kernel(const int64_t* __restrict__ idx, long long stride,
                                                                 const float* __restrict__ g, long long ldg, int V, int C,
                                                                 long long M, int hot_rows, float* __restrict__ dtable) {
  extern __shared__ float hot[];          // [hot_rows <= EMB_HOT][128]
  __shared__ unsigned long long used;     // bit v: row v < EMB_HOT was touched by this CTA
  const int c0 = blockIdx.y * 128, c = c0 + threadIdx.x;
  for (int v = 0; v < hot_rows; ++v) hot[v * 128 + threadIdx.x] = 0.f;
  const long long chunk = (M + gridDim.x - 1) / gridDim.x;
  const long long beg = blockIdx.x * chunk, end = (beg + chunk < M) ? beg + chunk : M;
  unsigned long long mine = 0ull;
  if (c < C) {
    long long m = beg;
    for (; m + 4 <= end; m += 4) {   // four rows in flight; the private-table updates stay in row order
      long long v[4];
      float x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = idx[(m + u) * stride];
        x[u] = __ldg(g + (m + u) * ldg + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (v[u] >= 0 && v[u] < V) {
          if (v[u] < hot_rows) {
            hot[v[u] * 128 + threadIdx.x] += x[u];
            mine |= 1ull << v[u];
          } else {
            atomicAdd(dtable + v[u] * C + c, x[u]);
          }
        }
      }
    }
    for (; m < end; ++m) {
      const long long v = idx[m * stride];
      if (v >= 0 && v < V) {
        const float x = __ldg(g + m * ldg + c);
        if (v < hot_rows) {
          hot[v * 128 + threadIdx.x] += x;
          mine |= 1ull << v;
        } else {
          atomicAdd(dtable + v * C + c, x);
        }
      }
    }
  }
  if (threadIdx.x == 0) used = mine;   // every thread of the CTA saw the same index sequence
  __syncthreads();
  const unsigned long long u = used;
  if (c < C)
    for (int v = 0; v < hot_rows; ++v)
      if ((u >> v) & 1ull) atomicAdd(dtable + (long long)v * C + c, hot[v * 128 + threadIdx.x]);
}
extern "C" int64_t sb_embedding_bwd_workspace_floats(int32_t V, int32_t C) {
  return (int64_t)sb_num_sms() * V * C;
}
extern "C" int sb_embedding_bwd(const int64_t* idx, int64_t stride, const float* g, int64_t ldg, int32_t V, int32_t C,
                                int64_t M, float* dtable, float* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = (int)sb_ceil_div(C, 32) * 32;
  SB_CHECK_ARG(threads <= 1024, "sb_embedding_bwd: feature dim too wide");
  const size_t smem = (size_t)V * threads * sizeof(float);
  if (M == 0 || smem > 200 * 1024) {
    SB_CUDA(cudaMemsetAsync(dtable, 0, sizeof(float) * V * C, st));
    if (M == 0) return SB_OK;
    long long blocks = sb_ceil_div(M, 32);            // >= 32 rows per CTA, up to 8 CTAs per SM
    const long long cap = (long long)sb_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const int hot_rows = V < EMB_HOT ? V : EMB_HOT;
    const size_t hot_smem = (size_t)hot_rows * 128 * sizeof(float);   // <= 32 KB
    dim3 grid((unsigned)blocks, (unsigned)sb_ceil_div(C, 128));
    embedding_bwd_hot_kernel<<<grid, 128, hot_smem, st>>>(idx, stride, g, ldg, V, C, M, hot_rows, dtable);
    SB_CHECK_LAUNCH("sb_embedding_bwd(atomic)");
    return SB_OK;
  }
  int grid = sb_num_sms();
  if (grid > M) grid = (int)M;
  SB_CUDA(cudaFuncSetAttribute(embedding_bwd_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  embedding_bwd_partial_kernel<<<grid, threads, smem, st>>>(idx, stride, g, ldg, V, C, M, workspace);
  SB_CHECK_LAUNCH("sb_embedding_bwd(partial)");
  embedding_bwd_reduce_kernel<<<(unsigned)sb_ceil_div((long long)V * C, 256), 256, 0, st>>>(workspace, grid, V, C,
                                                                                         dtable);
  SB_CHECK_LAUNCH("sb_embedding_bwd(reduce)");
  return SB_OK;
}

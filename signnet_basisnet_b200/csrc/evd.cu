// Batched Laplacian eigendecomposition on the device (the step BEFORE the hot path).
// The reference recomputes torch.linalg.eigh of every graph's normalised Laplacian on the CPU inside the DataLoader,
// every epoch (Alchemy/sign_net/transform.py:7-23, `transform=` not `pre_transform=` in main_alchemy.py:71-72), and the
// DGL tree once per dataset (GraphPrediction/data/molecules.py:148-181).  Here one WARP per graph builds
//     L = I - D^-1/2 A D^-1/2     (A symmetrised, de-duplicated, self loops removed; isolated nodes: D^-1/2 := 0;
//                                  = get_laplacian(to_undirected(edge_index), 'sym'), transform.py:18-20)
// in shared memory from the CSR and diagonalises it with the parallel cyclic Jacobi method: a round-robin schedule
// gives n/2 disjoint (p, q) pairs per step, lanes compute the rotations, then all lanes apply them to the columns of A
// and V and to the rows of A.  n <= 64 (molecules: n <= 37).  Output in the reference's layout: eigen_values[N] ascending
// per graph, eigen_vectors[sum n_b^2] row-major V[node, eig] (transform.py:14).  Eigenvector signs / bases inside
// degenerate eigenspaces are as arbitrary as LAPACK's; SignNet is invariant to the former by construction.
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define EVD_NMAX 64
#define EVD_WARPS 4
#define EVD_SWEEPS 14

__global__ void __launch_bounds__(32 * EVD_WARPS) laplacian_evd_kernel(const int32_t* __restrict__ graph_ptr,
                                                                       const int32_t* __restrict__ in_ptr,
                                                                       const int32_t* __restrict__ in_src,
                                                                       const int64_t* __restrict__ vec_ptr, int B, int nrows,
                                                                       int ld,
                                                                       float* __restrict__ evals,
                                                                       float* __restrict__ evecs, int* __restrict__ flags) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t per_warp = (size_t)2 * nrows * ld + 4 * 32 + 2 * EVD_NMAX;   // nrows = n_max of the batch
  float* A = sm + warp * per_warp;          // [n][ld]
  float* V = A + (size_t)nrows * ld;        // [n][ld]
  float* rc = V + (size_t)nrows * ld;       // rotation cos per pair
  float* rs = rc + 32;                      // rotation sin per pair
  int* rp = reinterpret_cast<int*>(rs + 32);
  int* rq = rp + 32;
  float* lam = reinterpret_cast<float*>(rq + 32);   // eigenvalues
  int* rank = reinterpret_cast<int*>(lam + EVD_NMAX);

  for (int b = blockIdx.x * EVD_WARPS + warp; b < B; b += gridDim.x * EVD_WARPS) {
    const int node0 = graph_ptr[b];
    const int n = graph_ptr[b + 1] - node0;
    if (n <= 0) continue;
    if (n > nrows) {
      if (lane == 0) atomicOr(flags, 1);
      continue;
    }
    // ---- adjacency (symmetrised, 0/1) and V = I
    for (int idx = lane; idx < n * n; idx += 32) {
      const int i = idx / n, j = idx - i * n;
      A[i * ld + j] = 0.f;
      V[i * ld + j] = (i == j) ? 1.f : 0.f;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      for (int e = in_ptr[node0 + i]; e < in_ptr[node0 + i + 1]; ++e) {
        const int j = in_src[e] - node0;
        if (j != i && j >= 0 && j < n) {
          A[i * ld + j] = 1.f;   // benign race: every writer stores 1
          A[j * ld + i] = 1.f;
        }
      }
    }
    __syncwarp();
    // degree -> D^-1/2 (kept in lam), then L
    for (int i = lane; i < n; i += 32) {
      float d = 0.f;
      for (int j = 0; j < n; ++j) d += A[i * ld + j];
      lam[i] = d > 0.f ? 1.0f / sqrtf(d) : 0.f;
    }
    __syncwarp();
    for (int idx = lane; idx < n * n; idx += 32) {
      const int i = idx / n, j = idx - i * n;
      const float a = A[i * ld + j];
      A[i * ld + j] = (i == j ? 1.f : 0.f) - lam[i] * a * lam[j];
    }
    __syncwarp();

    // ---- parallel cyclic Jacobi
    const int m = (n + 1) & ~1;          // players of the round-robin tournament (one dummy if n is odd)
    const int half = m >> 1;
    for (int sweep = 0; sweep < EVD_SWEEPS; ++sweep) {
      float off = 0.f;                    // sum of squares of the off-diagonal entries zeroed in this sweep
      for (int r = 0; r < m - 1; ++r) {
        if (lane < half) {
          int p, q;
          if (lane == 0) { p = m - 1; q = r; }
          else { p = (r + lane) % (m - 1); q = (r - lane + (m - 1)) % (m - 1); }
          if (p > q) { const int t = p; p = q; q = t; }
          float c = 1.f, s = 0.f;
          if (q < n) {
            const float apq = A[p * ld + q];
            off = fmaf(apq, apq, off);
            if (fabsf(apq) > 1e-30f) {
              const float tau = (A[q * ld + q] - A[p * ld + p]) / (2.f * apq);
              const float t = (tau >= 0.f ? 1.f : -1.f) / (fabsf(tau) + sqrtf(1.f + tau * tau));
              c = 1.0f / sqrtf(1.f + t * t);
              s = t * c;
            }
          } else {
            p = -1;   // pair with the dummy player: idle
          }
          rp[lane] = p; rq[lane] = q; rc[lane] = c; rs[lane] = s;
        }
        __syncwarp();
        // columns p, q of A and V:  X <- X J
        for (int idx = lane; idx < half * n; idx += 32) {
          const int pr = idx / n, k = idx - pr * n;
          const int p = rp[pr];
          if (p < 0) continue;
          const int q = rq[pr];
          const float c = rc[pr], s = rs[pr];
          const float ap = A[k * ld + p], aq = A[k * ld + q];
          A[k * ld + p] = c * ap - s * aq;
          A[k * ld + q] = s * ap + c * aq;
          const float vp = V[k * ld + p], vq = V[k * ld + q];
          V[k * ld + p] = c * vp - s * vq;
          V[k * ld + q] = s * vp + c * vq;
        }
        __syncwarp();
        // rows p, q of A:  A <- J^T A
        for (int idx = lane; idx < half * n; idx += 32) {
          const int pr = idx / n, k = idx - pr * n;
          const int p = rp[pr];
          if (p < 0) continue;
          const int q = rq[pr];
          const float c = rc[pr], s = rs[pr];
          const float ap = A[p * ld + k], aq = A[q * ld + k];
          A[p * ld + k] = c * ap - s * aq;
          A[q * ld + k] = s * ap + c * aq;
        }
        __syncwarp();
      }
      off = warp_sum(off);
      if (off < 3e-11f) break;            // entries of L are O(1): off-diagonal mass at fp32 resolution (~n^2/2 * eps^2)
    }

    // ---- ascending order (torch.linalg.eigh) and the reference's layout
    for (int j = lane; j < n; j += 32) lam[j] = A[j * ld + j];
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
      const float lj = lam[j];
      int rk = 0;
      for (int i = 0; i < n; ++i) {
        const float li = lam[i];
        rk += (li < lj || (li == lj && i < j)) ? 1 : 0;
      }
      rank[j] = rk;
      evals[node0 + rk] = lj;
    }
    __syncwarp();
    float* out = evecs + vec_ptr[b];
    for (int idx = lane; idx < n * n; idx += 32) {
      const int k = idx / n, j = idx - k * n;
      out[(size_t)k * n + rank[j]] = V[k * ld + j];
    }
    __syncwarp();
  }
}

extern "C" int sb_laplacian_evd(const int32_t* graph_ptr, const int32_t* in_ptr, const int32_t* in_src,
                                const int64_t* vec_ptr, int32_t B, int32_t nmax, float* eigen_values,
                                float* eigen_vectors, int32_t* flags, void* stream) {
  SB_CHECK_ARG(B >= 0 && nmax >= 0, "sb_laplacian_evd: bad sizes");
  SB_CHECK_ARG(nmax <= EVD_NMAX, "sb_laplacian_evd: graphs with more than %d nodes are not supported (n_max = %d)",
               EVD_NMAX, nmax);
  if (B == 0 || nmax == 0) return SB_OK;
  const int ld = (nmax | 1);   // odd row stride: column walks are bank-conflict free
  const size_t per_warp = ((size_t)2 * nmax * ld + 4 * 32 + 2 * EVD_NMAX) * sizeof(float);
  const size_t smem = per_warp * EVD_WARPS;
  static size_t configured = 0;
  if (smem > configured) {
    SB_CUDA(cudaFuncSetAttribute(laplacian_evd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int grid = (B + EVD_WARPS - 1) / EVD_WARPS;
  const int cap = sb_num_sms() * 8;
  if (grid > cap) grid = cap;
  laplacian_evd_kernel<<<grid, 32 * EVD_WARPS, smem, (cudaStream_t)stream>>>(graph_ptr, in_ptr, in_src, vec_ptr, B, nmax, ld,
                                                                            eigen_values, eigen_vectors, flags);
  SB_CHECK_LAUNCH("sb_laplacian_evd");
  return SB_OK;
}

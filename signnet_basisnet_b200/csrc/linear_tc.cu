// K2 on the 5th-generation tensor cores: y = f(x) W^T (+ bias, relu, column statistics) with tcgen05.mma kind::tf32
// and error-compensated operands (3xTF32):  x = xh + xl, w = wh + wl with xh/wh the tf32-rounded heads, and
//     x w  ~=  xh wh + xh wl + xl wh        (fp32 accumulation in tensor memory)
// which keeps the result within ~1e-6 relative of an fp32 FFMA contraction — BASELINE.json's 1e-5 parity bar rules out
// a single-pass TF32/bf16 product.  Same contract as linear_fwd_kernel (linear.cu): prologue none | BN-affine |
// BN-affine+ReLU applied once per element, epilogue bias / ReLU / per-(group, channel) fp64 column sums.
//
// One persistent CTA per SM, warp-specialised:
//   warps 0..15  workers: whole-tile register prefetch of the NEXT 128-row tile (64 KB of HBM reads in flight per SM),
//                prologue + head/tail split + swizzled st.shared of one 32-float K-block at a time into a 2-stage ring,
//                then the epilogue: tcgen05.ld (one 32x32 block per warp) -> bias/ReLU -> shared staging (the idle
//                ring) -> coalesced 128-bit stores + fp64 column sums for the BatchNorm that follows;
//   warp 16      MMA issuer: waits full[stage], issues 12 x UMMA 128 x N x 8 (kind::tf32) per K-block against the
//                resident weight (head and tail, canonical K-major SWIZZLE_128B, 128 KB), tcgen05.commit recycles the
//                ring stage (mma_done) and publishes the accumulator (acc_done).
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define TC_BM 128
#define TC_KB 32                      // floats per K-block = one 128-byte swizzle row
#define TC_BLK_BYTES (128 * 128)      // one [128 rows x 32 floats] operand block
#define TC_WORKERS 512
#define TC_THREADS (TC_WORKERS + 32)
#define TC_MAXG 2

struct TcArgs {
  const float* x;
  long long ldx;
  const float* w;
  long long w_rs, w_cs;
  const float* bias;
  float* y;
  long long ldy;
  long long R;
  int G, K, N, nkb, NP;
  int pro;
  const float* pa;
  const float* pc;
  int relu;
  double* stats;
  int accumulate;
  int xvec, yvec, ycols;
};

__device__ __forceinline__ uint64_t tc_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                          // LBO: unused for swizzled K-major operands
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;     // SBO: 1024 B between 8-row groups
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t tc_sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
// head = x rounded to nearest tf32 (10-bit mantissa), tail = x - head (exact in fp32; the tensor core keeps its top
// 11 bits).  Round-to-nearest instead of truncation removes the systematic sign-correlated bias of the heads.
__device__ __forceinline__ void tc_split(float x, float& h, float& l) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  h = __uint_as_float(u);
  l = x - h;
}
__device__ __forceinline__ void tc_mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_WORKERS) : "memory"); }
#define TC_LD32(v, taddr)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26," \
               "%27,%28,%29,%30,%31}, [%32];"                                                                         \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),           \
                 "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),           \
                 "=r"(v[30]), "=r"(v[31])                                                                            \
               : "r"(taddr))

__global__ void __launch_bounds__(TC_THREADS, 1) linear_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = a.nkb;
  uint8_t* Wh = smem;                              // [nkb][16 KB]
  uint8_t* Wl = Wh + nkb * TC_BLK_BYTES;
  uint8_t* ring = Wl + nkb * TC_BLK_BYTES;         // 2 stages x (head 16 KB | tail 16 KB) = 64 KB; reused as the
                                                   // [128][128] fp32 staging tile of the epilogue
  __shared__ double sacc[TC_MAXG * 2 * 128];
  __shared__ float s_pa[TC_MAXG * 128], s_pc[TC_MAXG * 128], s_bias[128];
  __shared__ uint64_t full[2], mma_done[2], acc_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, N = a.N;

  // ---- one-time: split the weight into head/tail, canonical K-major SW128 blocks; rows >= N and columns >= K are 0
  for (int idx = tid; idx < nkb * 128 * TC_KB; idx += TC_THREADS) {
    int n, k;
    if (a.w_cs == 1) { n = idx / (nkb * TC_KB); k = idx - n * (nkb * TC_KB); }
    else { k = idx / 128; n = idx - k * 128; }
    float v = 0.f;
    if (k < K && n < N) v = __ldg(a.w + (long long)n * a.w_rs + (long long)k * a.w_cs);
    float h, l;
    tc_split(v, h, l);
    const uint32_t off = (uint32_t)(k / TC_KB) * TC_BLK_BYTES + tc_sw128(n, (k % TC_KB) >> 2) + (uint32_t)(k & 3) * 4;
    *reinterpret_cast<float*>(Wh + off) = h;
    *reinterpret_cast<float*>(Wl + off) = l;
  }
  for (int idx = tid; idx < TC_MAXG * 128; idx += TC_THREADS) {
    const int g = idx >> 7, c = idx & 127;
    const bool ok = a.pro && g < a.G && c < K;
    s_pa[idx] = ok ? __ldg(a.pa + (long long)g * K + c) : 1.f;
    s_pc[idx] = ok ? __ldg(a.pc + (long long)g * K + c) : 0.f;
  }
  for (int idx = tid; idx < 128; idx += TC_THREADS) s_bias[idx] = (a.bias && idx < N) ? __ldg(a.bias + idx) : 0.f;
  for (int idx = tid; idx < TC_MAXG * 2 * 128; idx += TC_THREADS) sacc[idx] = 0.0;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_init(&acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const long long tpg = (a.R + TC_BM - 1) / TC_BM;
  const long long ntiles = tpg * a.G;

  if (warp == 16) {
    // =============================================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.NP >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      unsigned cnt = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int stage = cnt & 1;
          mbar_wait(&full[stage], (cnt >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = smem_u32(ring + stage * 2 * TC_BLK_BYTES), al = ah + TC_BLK_BYTES;
          const uint32_t wh = smem_u32(Wh + kb * TC_BLK_BYTES), wl = smem_u32(Wl + kb * TC_BLK_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t o = j * 32;
            tc_mma(tmem, tc_make_desc(ah + o), tc_make_desc(wh + o), idesc, (kb | j) ? 1u : 0u);
            tc_mma(tmem, tc_make_desc(ah + o), tc_make_desc(wl + o), idesc, 1u);
            tc_mma(tmem, tc_make_desc(al + o), tc_make_desc(wh + o), idesc, 1u);
          }
          tc_commit(&mma_done[stage]);
          if (kb == nkb - 1) tc_commit(&acc_done);
        }
      }
    }
  } else {
    // ================================================================================================== workers
    // Whole-tile register prefetch (raw loads only: nothing here may depend on the loaded values).
    float4 pre[8];
    auto load_tile = [&](long long tile) {
      const int g = (int)(tile / tpg);
      const long long row0 = (tile - (long long)g * tpg) * TC_BM;
      const int rows = (int)((a.R - row0 < TC_BM) ? (a.R - row0) : TC_BM);
      const long long base = (long long)g * a.R + row0;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        if (kb >= nkb) break;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int f = tid + TC_WORKERS * q;
          const int row = f >> 3, c4 = f & 7;
          const int col = kb * TC_KB + c4 * 4;
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
          }
          pre[kb * 2 + q] = v;
        }
      }
    };
    // prologue (BatchNorm affine / ReLU) + zeroing of the K tail + head/tail split, applied when a block is consumed
    auto store_block = [&](int stage, int kb, int g, int rows) {
      uint8_t* sh = ring + stage * 2 * TC_BLK_BYTES;
      uint8_t* sl = sh + TC_BLK_BYTES;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int f = tid + TC_WORKERS * q;
        const int row = f >> 3, c4 = f & 7;
        const int col = kb * TC_KB + c4 * 4;
        const float4 pv = (kb == 0) ? pre[q] : (kb == 1) ? pre[2 + q] : (kb == 2) ? pre[4 + q] : pre[6 + q];
        float t[4] = {pv.x, pv.y, pv.z, pv.w};
        const bool live = row < rows;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (live && col + j < K) {
            if (a.pro) {
              const float u = fmaf(s_pa[g * 128 + col + j], t[j], s_pc[g * 128 + col + j]);
              t[j] = (a.pro == 2) ? fmaxf(u, 0.f) : u;
            }
          } else {
            t[j] = 0.f;
          }
        }
        float4 h, l;
        tc_split(t[0], h.x, l.x);
        tc_split(t[1], h.y, l.y);
        tc_split(t[2], h.z, l.z);
        tc_split(t[3], h.w, l.w);
        const uint32_t off = tc_sw128(row, c4);
        *reinterpret_cast<float4*>(sh + off) = h;
        *reinterpret_cast<float4*>(sl + off) = l;
      }
    };

    long long tile = blockIdx.x;
    unsigned cnt = 0, tile_it = 0;
    if (tile < ntiles) load_tile(tile);

    while (tile < ntiles) {
      const int g = (int)(tile / tpg);
      const long long row0 = (tile - (long long)g * tpg) * TC_BM;
      const int rows = (int)((a.R - row0 < TC_BM) ? (a.R - row0) : TC_BM);
      const long long base = (long long)g * a.R + row0;
      const long long ntile = tile + gridDim.x;

      for (int kb = 0; kb < nkb; ++kb, ++cnt) {
        const int stage = cnt & 1;
        const unsigned use = cnt >> 1;
        if (use > 0) mbar_wait(&mma_done[stage], (use - 1) & 1);   // MMAs that read this stage have retired
        store_block(stage, kb, g, rows);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        worker_sync();
        if (tid == 0) mbar_arrive(&full[stage]);
      }
      // registers of this tile are consumed: request the whole next tile while the tensor core + epilogue run
      if (ntile < ntiles) load_tile(ntile);

      // ---------------------------------------------------------------------------------------------- epilogue
      mbar_wait(&acc_done, tile_it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* stg = reinterpret_cast<float*>(ring);   // [128][128] fp32, float4 chunk index XOR (row & 31)
      {
        const int q = warp & 3, cb = warp >> 2;       // TMEM lane quarter, 32-column block
        const int row = q * 32 + lane;
        const int c0 = cb * 32;
        if (c0 < a.NP) {
          uint32_t v[32];
          TC_LD32(v, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = c0 + i * 4 + j;
              float t = __uint_as_float(v[i * 4 + j]) + s_bias[col];
              if (a.accumulate && row < rows && col < N) t += a.y[(base + row) * a.ldy + col];
              if (a.relu) t = fmaxf(t, 0.f);
              o[j] = (col < N) ? t : 0.f;
            }
            const int ch = (c0 >> 2) + i;
            *reinterpret_cast<float4*>(stg + row * 128 + ((ch ^ (row & 31)) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      worker_sync();
      // coalesced write-back: one warp per row (8 rows per warp), lanes over float4 chunks
      {
        const int col0 = lane * 4;
        if (col0 < a.ycols) {
#pragma unroll 4
          for (int row = warp; row < rows; row += TC_WORKERS / 32) {
            const float4 o = *reinterpret_cast<const float4*>(stg + row * 128 + ((lane ^ (row & 31)) << 2));
            float* yrow = a.y + (base + row) * a.ldy;
            if (a.yvec) {
              *reinterpret_cast<float4*>(yrow + col0) = o;
            } else {
              const float t[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col0 + j < a.ycols) yrow[col0 + j] = t[j];
            }
          }
        }
      }
      if (a.stats) {
        // column sums in fp64: thread -> (column, row quarter); two independent chains per thread
        const int col = tid & 127, part = tid >> 7;
        if (col < N) {
          double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
          const int r_beg = part * 32;
          const int r_end = (rows < r_beg + 32) ? rows : r_beg + 32;
          int row = r_beg;
          for (; row + 1 < r_end; row += 2) {
            const float t0 = stg[row * 128 + ((((col >> 2) ^ (row & 31)) << 2) | (col & 3))];
            const float t1 = stg[(row + 1) * 128 + ((((col >> 2) ^ ((row + 1) & 31)) << 2) | (col & 3))];
            s0 += (double)t0; q0 += (double)t0 * (double)t0;
            s1 += (double)t1; q1 += (double)t1 * (double)t1;
          }
          if (row < r_end) {
            const float t0 = stg[row * 128 + ((((col >> 2) ^ (row & 31)) << 2) | (col & 3))];
            s0 += (double)t0; q0 += (double)t0 * (double)t0;
          }
          atomicAdd(&sacc[(g * 2 + 0) * 128 + col], s0 + s1);
          atomicAdd(&sacc[(g * 2 + 1) * 128 + col], q0 + q1);
        }
      }
      worker_sync();   // staging (= ring) is free again; the TMEM accumulator has been drained
      tile = ntile;
      ++tile_it;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (a.stats) {
    for (int idx = tid; idx < a.G * 2 * 128; idx += TC_THREADS) {
      const int col = idx & 127, gj = idx >> 7;
      if (col < N && sacc[idx] != 0.0) atomicAdd(a.stats + (long long)gj * N + col, sacc[idx]);
    }
  }
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// Returns SB_ERR_UNSUPPORTED (without setting an error) when the shape is better served by the FFMA kernel.
int sb_linear_tc_launch(const float* x, int64_t ldx, const float* w, int64_t w_rs, int64_t w_cs, const float* bias,
                        float* y, int64_t ldy, int64_t R, int32_t G, int32_t K, int32_t N, int32_t pro,
                        const float* pa, const float* pc, int32_t relu, double* stats, int32_t accumulate,
                        int32_t ycols, cudaStream_t st) {
  if (K < 16 || K > 128 || N < 16 || N > 128 || G > TC_MAXG || R * G < 4096) return SB_ERR_UNSUPPORTED;
  TcArgs a;
  a.x = x; a.ldx = ldx; a.w = w; a.w_rs = w_rs; a.w_cs = w_cs; a.bias = bias; a.y = y; a.ldy = ldy; a.R = R; a.G = G;
  a.K = K; a.N = N; a.nkb = (K + TC_KB - 1) / TC_KB; a.NP = (N + 15) / 16 * 16;
  a.pro = pro; a.pa = pa; a.pc = pc; a.relu = relu; a.stats = stats; a.accumulate = accumulate;
  a.xvec = (ldx % 4 == 0) && ((uintptr_t)x % 16 == 0);
  a.yvec = (ldy % 4 == 0) && ((uintptr_t)y % 16 == 0);
  a.ycols = (ycols < a.NP) ? ycols : a.NP;
  const size_t smem = (size_t)2 * a.nkb * TC_BLK_BYTES + 4 * TC_BLK_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    SB_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 2 * 4 * TC_BLK_BYTES + 4 * TC_BLK_BYTES + 1024));
    configured = true;
  }
  const long long ntiles = sb_ceil_div(R, TC_BM) * G;
  long long grid = sb_num_sms();
  if (grid > ntiles) grid = ntiles;
  linear_tc_kernel<<<(unsigned)grid, TC_THREADS, smem, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_fwd(tcgen05)");
  return SB_OK;
}

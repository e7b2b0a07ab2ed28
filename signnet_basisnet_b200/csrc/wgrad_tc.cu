// Weight gradient on the 5th-generation tensor cores:  dW[n, k] = sum_rows g[row, n] * f(x[row, k])  (+ dbias) with
// tcgen05.mma kind::tf32 and 3xTF32 error compensation (see linear_tc.cu).  The reduction dimension (rows) is the MMA K
// dimension, so both operands are consumed MN-major — exactly their natural row-major [row][feature] storage; for
// tf32 the only MN-major shared-memory layout is SWIZZLE_128B_BASE32B (4-row x 128-byte atoms, 32-byte XOR swizzle).
//
// Persistent CTA per SM, 16 worker warps + 1 MMA warp.  Workers stream 32-row chunks of g and x (register prefetch two
// chunks ahead, backed by bulk L2 prefetches WT_L2_AHEAD chunks ahead: with registers alone only 64 KB per SM were in
// flight and the kernel sat at 22 % of DRAM bandwidth stalled on the loads, profiles/r1g_wgrad_tc_ncu.csv), apply the forward prologue to x (BatchNorm affine / ReLU recomputed, never stored), split head/tail
// and write the four operand blocks of a 3-stage ring; the MMA warp issues 12 x UMMA 128 x K x 8 per chunk into one
// [128 x 128] fp32 accumulator in tensor memory that lives for the whole kernel; per-CTA partials are then added in a
// fixed order by wgrad_reduce_kernel (deterministic, no float atomics).
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define WT_ROWS 32
#define WT_BLK_BYTES (32 * 128 * 4)   // one [32 rows x 128 features] fp32 operand block = 16 KB
#define WT_STAGE_BYTES (4 * WT_BLK_BYTES)
#define WT_STAGES 3
#define WT_WORKERS 512
#define WT_THREADS (WT_WORKERS + 32)
#define WT_MAXG 2
#define WT_SLOTS 3       // register prefetch depth in chunks (3 x 32 KB in flight per SM on top of the L2 prefetch)
#define WT_L2_AHEAD 8   // chunks requested into L2 ahead of the register prefetch (cp.async.bulk.prefetch.L2)

struct WgTcArgs {
  const float* g;
  long long ldg;
  const float* x;
  long long ldx;
  long long R;
  int G, N, K, KP;
  int pro;
  const float* pa;
  const float* pc;
  float* part_w;   // [grid][128][128]
  float* part_b;   // [grid][128] or null
  int gvec, xvec;
};

__device__ __forceinline__ uint64_t wt_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((4096 >> 4) & 0x3FFF) << 16;   // LBO: next 32-feature MN block
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;    // SBO: next atom of 4 rows
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)1 << 61;                        // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void wt_split(float x, float& h, float& l) {
  // == cvt.rna.tf32.f32 for finite x, in two ALU instructions (see tc_split in linear_tc.cu)
  h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  l = x - h;
}
__device__ __forceinline__ void wt_mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void wt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wt_worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WT_WORKERS) : "memory"); }

__device__ __forceinline__ void wt_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// FAST: N == K == 128 and 16-byte aligned rows -> no per-element bounds / alignment predicates are compiled in.
template <bool FAST>
__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_tc_kernel(const WgTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;   // [WT_STAGES][g_hi | g_lo | x_hi | x_lo]
  __shared__ float s_pa[WT_MAXG * 128], s_pc[WT_MAXG * 128];
  __shared__ uint64_t full[WT_STAGES], mma_done[WT_STAGES], acc_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, K = a.K;

  for (int idx = tid; idx < WT_MAXG * 128; idx += WT_THREADS) {
    const int g = idx >> 7, c = idx & 127;
    const bool ok = a.pro && g < a.G && c < K;
    s_pa[idx] = ok ? __ldg(a.pa + (long long)g * K + c) : 1.f;
    s_pc[idx] = ok ? __ldg(a.pc + (long long)g * K + c) : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < WT_STAGES; ++s) {
      mbar_init(&full[s], WT_WORKERS / 32);
      mbar_init(&mma_done[s], 1);
    }
    mbar_init(&acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const long long cpg = (a.R + WT_ROWS - 1) / WT_ROWS;
  const long long nch = cpg * a.G;

  if (warp == 16) {
    // =============================================================================================== MMA issuer
    if (lane == 0) {
      // D[M = 128 (n), N = KP (k)] += A^T-major G chunk x X chunk, both MN-major tf32
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(a.KP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      unsigned cnt = 0;
      for (long long c = blockIdx.x; c < nch; c += gridDim.x, ++cnt) {
        const int stage = cnt % WT_STAGES;
        mbar_wait(&full[stage], (cnt / WT_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t gh = smem_u32(ring + stage * WT_STAGE_BYTES), gl = gh + WT_BLK_BYTES;
        const uint32_t xh = gl + WT_BLK_BYTES, xl = xh + WT_BLK_BYTES;
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // 4 groups of 8 rows
          const uint32_t o = j * 1024;
          wt_mma(tmem, wt_make_desc(gh + o), wt_make_desc(xh + o), idesc, (cnt | j) ? 1u : 0u);
          wt_mma(tmem, wt_make_desc(gh + o), wt_make_desc(xl + o), idesc, 1u);
          wt_mma(tmem, wt_make_desc(gl + o), wt_make_desc(xh + o), idesc, 1u);
        }
        wt_commit(&mma_done[stage]);
      }
      wt_commit(&acc_done);
    }
  } else {
    // ================================================================================================== workers
    // item q of a thread: idx = tid + 512 q -> 32-feature block (idx >> 8), row (idx >> 3) & 31, float4 (idx & 7)
    // WT_SLOTS register prefetch slots; the chunk loop is unrolled by WT_SLOTS so every slot index is a compile-time
    // constant: a chunk waits only for ITS loads (issued WT_SLOTS chunks earlier).  The first version selected the slot
    // with predicated moves, which made every chunk also wait for the loads issued one chunk earlier.
    float4 pg[WT_SLOTS][2], px[WT_SLOTS][2];
    float dbs[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const int row = (tid >> 3) & 31, c4 = tid & 7;
#define WT_LOAD(CC, S)                                                                         \
    {                                                                                          \
      const long long c_ = (CC);                                                               \
      const int g_ = (c_ >= cpg) ? 1 : 0; /* G <= WT_MAXG = 2: no 64-bit division */           \
      const long long row0_ = (c_ - (long long)g_ * cpg) * WT_ROWS;                            \
      const int rows_ = (int)((a.R - row0_ < WT_ROWS) ? (a.R - row0_) : WT_ROWS);              \
      const long long base_ = (long long)g_ * a.R + row0_;                                     \
      _Pragma("unroll") for (int q = 0; q < 2; ++q) {                                          \
        const int col = ((tid + WT_WORKERS * q) >> 8) * 32 + c4 * 4;                           \
        float4 vg = make_float4(0.f, 0.f, 0.f, 0.f), vx = vg;                                  \
        if (FAST) {                                                                            \
          if (row < rows_) {                                                                   \
            vg = ldg4(a.g + (base_ + row) * a.ldg + col);                                      \
            vx = ldg4(a.x + (base_ + row) * a.ldx + col);                                      \
          }                                                                                    \
        } else if (row < rows_) {                                                              \
          if (col < N) {                                                                       \
            const float* p = a.g + (base_ + row) * a.ldg + col;                                \
            if (a.gvec) vg = ldg4(p);                                                          \
            else {                                                                             \
              vg.x = __ldg(p);                                                                 \
              if (col + 1 < N) vg.y = __ldg(p + 1);                                            \
              if (col + 2 < N) vg.z = __ldg(p + 2);                                            \
              if (col + 3 < N) vg.w = __ldg(p + 3);                                            \
            }                                                                                  \
          }                                                                                    \
          if (col < K) {                                                                       \
            const float* p = a.x + (base_ + row) * a.ldx + col;                                \
            if (a.xvec) vx = ldg4(p);                                                          \
            else {                                                                             \
              vx.x = __ldg(p);                                                                 \
              if (col + 1 < K) vx.y = __ldg(p + 1);                                            \
              if (col + 2 < K) vx.z = __ldg(p + 2);                                            \
              if (col + 3 < K) vx.w = __ldg(p + 3);                                            \
            }                                                                                  \
          }                                                                                    \
        }                                                                                      \
        pg[S][q] = vg;                                                                         \
        px[S][q] = vx;                                                                         \
      }                                                                                        \
    }
    auto store_chunk = [&](int stage, const float4 (&vgs)[2], const float4 (&vxs)[2], long long c) {
      const int g = (c >= cpg) ? 1 : 0;
      const long long row0 = (c - (long long)g * cpg) * WT_ROWS;
      const int rows = (int)((a.R - row0 < WT_ROWS) ? (a.R - row0) : WT_ROWS);
      uint8_t* sb = ring + stage * WT_STAGE_BYTES;
      const bool live = row < rows;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int blk = (tid + WT_WORKERS * q) >> 8;
        const int col = blk * 32 + c4 * 4;
        float tg[4] = {vgs[q].x, vgs[q].y, vgs[q].z, vgs[q].w};
        float tx[4] = {vxs[q].x, vxs[q].y, vxs[q].z, vxs[q].w};
        if (a.pro) {   // coefficients of columns >= K are (1, 0): harmless, those entries are zeroed below
          const float4 pa4 = *reinterpret_cast<const float4*>(&s_pa[g * 128 + col]);
          const float4 pc4 = *reinterpret_cast<const float4*>(&s_pc[g * 128 + col]);
          tx[0] = fmaf(pa4.x, tx[0], pc4.x); tx[1] = fmaf(pa4.y, tx[1], pc4.y);
          tx[2] = fmaf(pa4.z, tx[2], pc4.z); tx[3] = fmaf(pa4.w, tx[3], pc4.w);
          if (a.pro == 2) {
            tx[0] = fmaxf(tx[0], 0.f); tx[1] = fmaxf(tx[1], 0.f); tx[2] = fmaxf(tx[2], 0.f); tx[3] = fmaxf(tx[3], 0.f);
          }
        }
        if (FAST) {
          if (!live) { tg[0] = tg[1] = tg[2] = tg[3] = 0.f; tx[0] = tx[1] = tx[2] = tx[3] = 0.f; }
        } else if (!(live && col + 3 < N && col + 3 < K)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!(live && col + j < N)) tg[j] = 0.f;
            if (!(live && col + j < K)) tx[j] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) dbs[q][j] += tg[j];
        float4 h, l;
        const uint32_t off = (uint32_t)(blk * 4096 + row * 128 + (((c4 >> 1) ^ (row & 3)) << 5) + (c4 & 1) * 16);
        wt_split(tg[0], h.x, l.x); wt_split(tg[1], h.y, l.y); wt_split(tg[2], h.z, l.z); wt_split(tg[3], h.w, l.w);
        *reinterpret_cast<float4*>(sb + off) = h;
        *reinterpret_cast<float4*>(sb + WT_BLK_BYTES + off) = l;
        wt_split(tx[0], h.x, l.x); wt_split(tx[1], h.y, l.y); wt_split(tx[2], h.z, l.z); wt_split(tx[3], h.w, l.w);
        *reinterpret_cast<float4*>(sb + 2 * WT_BLK_BYTES + off) = h;
        *reinterpret_cast<float4*>(sb + 3 * WT_BLK_BYTES + off) = l;
      }
    };
    // one thread asks L2 for a whole chunk (rows are contiguous: 32 x ld floats per operand)
    const bool l2ok = FAST || (a.gvec && a.xvec);
    auto l2_chunk = [&](long long cc) {
      const int g = (cc >= cpg) ? 1 : 0;
      const long long row0 = (cc - (long long)g * cpg) * WT_ROWS;
      const int rows = (int)((a.R - row0 < WT_ROWS) ? (a.R - row0) : WT_ROWS);
      const long long base = (long long)g * a.R + row0;
      wt_prefetch_l2(a.g + base * a.ldg, (uint32_t)(rows * a.ldg * 4));
      wt_prefetch_l2(a.x + base * a.ldx, (uint32_t)(rows * a.ldx * 4));
    };
    long long c = blockIdx.x;
    unsigned cnt = 0;
    if (tid == 32 && l2ok)
      for (int s = WT_SLOTS; s < WT_L2_AHEAD; ++s)
        if (c + (long long)s * gridDim.x < nch) l2_chunk(c + (long long)s * gridDim.x);
#pragma unroll
    for (int s = 0; s < WT_SLOTS; ++s)
      if (c + (long long)s * gridDim.x < nch) WT_LOAD(c + (long long)s * gridDim.x, s)
    while (c < nch) {
#pragma unroll
      for (int s = 0; s < WT_SLOTS; ++s) {
        if (c < nch) {   // CTA-uniform
          const int stage = cnt % WT_STAGES;
          const unsigned use = cnt / WT_STAGES;
          if (use > 0) mbar_wait(&mma_done[stage], (use - 1) & 1);
          store_chunk(stage, pg[s], px[s], c);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[stage]);   // one arrival per worker warp: no block-wide barrier per chunk
          if (tid == 32 && l2ok) {
            const long long cl = c + (long long)WT_L2_AHEAD * gridDim.x;
            if (cl < nch) l2_chunk(cl);
          }
          const long long cn = c + (long long)WT_SLOTS * gridDim.x;   // refill the slot just consumed
          if (cn < nch) WT_LOAD(cn, s)
          c += gridDim.x;
          ++cnt;
        }
      }
    }
#undef WT_LOAD

    // ---- epilogue: accumulator -> per-CTA partial (row n per thread, 32 columns per warp)
    mbar_wait(&acc_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      const int q = warp & 3, cb = warp >> 2;
      const int row = q * 32 + lane, c0 = cb * 32;
      float* dst = a.part_w + ((size_t)blockIdx.x * 128 + row) * 128 + c0;
      if (c0 < a.KP) {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
                     "%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(dst + i * 4) = make_float4(__uint_as_float(v[i * 4]), __uint_as_float(v[i * 4 + 1]),
                                                                __uint_as_float(v[i * 4 + 2]), __uint_as_float(v[i * 4 + 3]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(dst + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // ---- dbias: fixed-order reduction of the per-thread column partials through shared memory (ring is idle now)
    if (a.part_b) {
      wt_worker_sync();
      float* red = reinterpret_cast<float*>(ring);   // [32 row slots][128 columns]
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int idx = tid + WT_WORKERS * q;
        const int rslot = (idx >> 3) & 31, col = (idx >> 8) * 32 + (idx & 7) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) red[rslot * 128 + col + j] = dbs[q][j];
      }
      wt_worker_sync();
      if (tid < 128) {
        float s = 0.f;
        for (int r = 0; r < 32; ++r) s += red[r * 128 + tid];
        a.part_b[(size_t)blockIdx.x * 128 + tid] = s;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int sb_wgrad_reduce_launch(const float* part_w, const float* part_b, int nparts, int BN, int BK, int N, int K, float* dw,
                           long long rs, long long cs, float* db, int accumulate, cudaStream_t st);

// Returns SB_ERR_UNSUPPORTED (without setting an error) when the shape is better served by the FFMA kernel.
int sb_wgrad_tc_launch(const float* gy, int64_t ldg, const float* x, int64_t ldx, int64_t R, int32_t G, int32_t N,
                       int32_t K, int32_t pro, const float* pa, const float* pc, float* dw, int64_t dw_rs,
                       int64_t dw_cs, float* db, int32_t accumulate, float* workspace, cudaStream_t st) {
  if (K < 16 || K > 128 || N < 16 || N > 128 || G > WT_MAXG || R * G < 4096) return SB_ERR_UNSUPPORTED;
  WgTcArgs a;
  a.g = gy; a.ldg = ldg; a.x = x; a.ldx = ldx; a.R = R; a.G = G; a.N = N; a.K = K; a.KP = (K + 15) / 16 * 16;
  a.pro = pro; a.pa = pa; a.pc = pc;
  a.gvec = (ldg % 4 == 0) && ((uintptr_t)gy % 16 == 0);
  a.xvec = (ldx % 4 == 0) && ((uintptr_t)x % 16 == 0);
  const long long nch = sb_ceil_div(R, WT_ROWS) * G;
  long long grid = sb_num_sms();
  if (grid > nch) grid = nch;
  a.part_w = workspace;
  a.part_b = db ? workspace + (size_t)grid * 128 * 128 : nullptr;
  const size_t smem = (size_t)WT_STAGES * WT_STAGE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    SB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if (a.gvec && a.xvec && N == 128 && K == 128) wgrad_tc_kernel<true><<<(unsigned)grid, WT_THREADS, smem, st>>>(a);
  else wgrad_tc_kernel<false><<<(unsigned)grid, WT_THREADS, smem, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_wgrad(tcgen05)");
  return sb_wgrad_reduce_launch(a.part_w, a.part_b, (int)grid, 128, 128, N, K, dw, dw_rs, dw_cs, db, accumulate, st);
}

// K2 on tcgen05, WEIGHT-STATIONARY IN TENSOR MEMORY (fourth generation of linear_tc.cu, fast shapes only).
//
// Why: scripts/pair_check.cu + ncu on the B200 (profiles/r1z_pair_check_mode*.log, r1z_linear_tma*_ncu_hot.txt) show that
// neither halving the LSU traffic (TMA-fed kernel, linear_tc_ws.cu) nor splitting the weight over a CTA pair
// (linear_tc_pair.cu) makes the Linear faster: in every variant the producer warps spend more than half of their
// samples waiting for input, with only 2 K-blocks (32 KB) per SM in flight, because the split weight (head + tail,
// 128 KB at K = N = 128) leaves shared memory for a 2-stage operand ring only.  This kernel swaps the operand roles:
//       y^T [N x rows] = W [N x K] * x^T            (UMMA: D[M = N_out, N = 128 tile rows] = A * B^T, both K-major)
//   * A = the split weight lives in TENSOR MEMORY for the whole kernel (head: K columns, tail: K columns; lane = output
//     channel), written once with tcgen05.st - it costs no shared memory at all;
//   * B = the activation K-block [128 rows x 32 floats], TMA-loaded raw into a 7-stage ring (224 KB: 7 K-blocks = 112 KB
//     of reads in flight per SM instead of 32 KB), tail computed in place by the producer warps as in linear_tc_ws.cu;
//     the tensor cores read 4 KB of shared memory per MMA instead of 8 KB;
//   * the accumulator comes out TRANSPOSED (lane = output channel, column = tile row), which is exactly what the epilogue
//     wants: a thread owns one channel, so bias, ReLU and the BatchNorm column sums are per-thread scalars (no shuffles,
//     no staging tile), and for each tile row the 32 lanes of a warp store 128 contiguous bytes (coalesced STG.32)
//     - no shared-memory transposition, no epilogue staging buffer.
// TMEM budget: 2 accumulators x 128 columns + weight head 128 + weight tail 128 = 512 columns (all of it).
//
// STATUS: opt-in (sb_set_tensor_cores(5) or SB_LINEAR_TMA=3); compiles for sm_100a, NOT yet run on a GPU (the round's
// GPU budget was spent when it was written) - first item of the next GPU visit: scripts/gpu_pair_check.sh 5.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define TW_BM 128
#define TW_KB 32
#define TW_BLK_BYTES (128 * 128)
#define TW_WORKERS 512
#define TW_PROD 256
#define TW_THREADS (TW_WORKERS + 64)   // + MMA warp (16) + TMA-issue warp (17)
#define TW_MAXG 2
#define TW_L2_AHEAD 2

struct TwArgs {
  const float* x;
  long long ldx;
  const float* w;
  long long w_rs, w_cs;
  const float* bias;
  float* y;
  long long ldy;
  long long R;
  int G, K, N, nkb;
  int pro;
  const float* pa;
  const float* pc;
  int relu;
  double* stats;
  int rawhead;
  int l2_ahead;   // tiles requested into L2 ahead of the TMA loads (SB_TMA_L2_AHEAD, default TW_L2_AHEAD)
};

__device__ __forceinline__ uint64_t tw_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t tw_sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void tw_split(float x, float& h, float& l) {   // == tc_split (linear_tc.cu)
  h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  l = x - h;
}
__device__ __forceinline__ float tw_tail_trunc(float x) {   // x minus the tf32 the tensor core reads from the raw word
  return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}
__device__ __forceinline__ void tw_mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tw_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 3-D tensor-map copies (SASS: UTMALDG / UTMASTG); coordinates = (column, row inside the group, group)
__device__ __forceinline__ void tw_tma_load(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tw_tma_store(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
#define TW_LD32(v, taddr)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26," \
               "%27,%28,%29,%30,%31}, [%32];"                                                                         \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),           \
                 "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),           \
                 "=r"(v[30]), "=r"(v[31])                                                                            \
               : "r"(taddr))


#define TW_STAGES 7                      // 7 x (head 16 KB | tail 16 KB) = 224 KB of shared memory
#define TW_COL_WH 256                    // tensor-memory columns: [0, 256) two accumulators, then weight head / tail
#define TW_COL_WL 384

// A operand from tensor memory (lane = row of A = output channel, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void tw_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
#define TW_ST32(taddr, v)                                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                       \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27," \
               "%28,%29,%30,%31,%32};"                                                                                \
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),   \
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),         \
                 "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),       \
                 "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])        \
               : "memory")

__global__ void __launch_bounds__(TW_THREADS, 1)
linear_tc_ws_kernel(const TwArgs a, const __grid_constant__ CUtensorMap tmx) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;                            // TW_STAGES x (head 16 KB | tail 16 KB)
  __shared__ __align__(16) float s_pa[TW_MAXG * 128], s_pc[TW_MAXG * 128];
  __shared__ uint64_t tma_full[TW_STAGES], full[TW_STAGES], mma_done[TW_STAGES], acc_done[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = a.nkb, K = a.K, N = a.N;

  for (int idx = tid; idx < TW_MAXG * 128; idx += TW_THREADS) {
    const int g = idx >> 7, c = idx & 127;
    const bool ok = a.pro && g < a.G && c < K;
    s_pa[idx] = ok ? __ldg(a.pa + (long long)g * K + c) : 1.f;
    s_pc[idx] = ok ? __ldg(a.pc + (long long)g * K + c) : 0.f;
  }
  if (tid == 0) {
    for (int i = 0; i < TW_STAGES; ++i) {
      mbar_init(&tma_full[i], 1);
      mbar_init(&full[i], TW_PROD / 32);
      mbar_init(&mma_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_done[i], 1);
      mbar_init(&acc_free[i], (TW_WORKERS - TW_PROD) / 32);
    }
    mbar_fence_init();
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  // ---- one-time: the split weight into tensor memory.  Warps 8..15 (lane quarter = warp & 3): warps 8..11 write the
  // heads, 12..15 the tails; thread -> output channel n = 32 q + lane, 32 K-columns per tcgen05.st.  Channels >= N and
  // columns >= K hold zeros (M is always 128).
  if (warp >= 8 && warp < 16) {
    const int q = warp & 3, is_tail = (warp >= 12) ? 1 : 0;
    const int n = q * 32 + lane;
    for (int c = 0; c < nkb; ++c) {
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int k = c * 32 + j;
        float wv = 0.f;
        if (n < N && k < K) wv = __ldg(a.w + (long long)n * a.w_rs + (long long)k * a.w_cs);
        float h, l;
        tw_split(wv, h, l);
        v[j] = __float_as_uint(is_tail ? l : h);
      }
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((is_tail ? TW_COL_WL : TW_COL_WH) + c * 32);
      TW_ST32(taddr, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const long long tpg = (a.R + TW_BM - 1) / TW_BM;
  const long long ntiles = tpg * a.G;

  if (warp == 16) {
    // =============================================================================================== MMA issuer
    if (lane == 0) {
      // D [M = 128 channels, N = 128 tile rows] += A (tensor memory, K-major) * B^T (shared memory, K-major)
      const uint32_t idesc =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TW_BM >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      unsigned cnt = 0, ti = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        const uint32_t buf = ti & 1u;
        if (ti >= 2) {
          mbar_wait(&acc_free[buf], ((ti >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tacc = tmem + buf * 128u;
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const unsigned stage = cnt % TW_STAGES;
          mbar_wait(&full[stage], (cnt / TW_STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t xh = smem_u32(ring + stage * 2 * TW_BLK_BYTES), xl = xh + TW_BLK_BYTES;
          const uint32_t wh = tmem + TW_COL_WH + kb * 32, wl = tmem + TW_COL_WL + kb * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tw_mma_ts(tacc, wh + j * 8, tw_make_desc(xh + j * 32), idesc, (kb | j) ? 1u : 0u);
            tw_mma_ts(tacc, wl + j * 8, tw_make_desc(xh + j * 32), idesc, 1u);
            tw_mma_ts(tacc, wh + j * 8, tw_make_desc(xl + j * 32), idesc, 1u);
          }
          tw_commit(&mma_done[stage]);
          if (kb == nkb - 1) tw_commit(&acc_done[buf]);
        }
      }
    }
  } else if (warp < TW_PROD / 32) {
    // ================================================================================================ producers
    // thread -> 16-byte chunk c4 of rows (tid >> 3) + 32 q, q < 4, of the K-block the TMA unit has put into the ring
    const int prow = tid >> 3, c4 = tid & 7;
    const bool raw = a.rawhead && !a.pro;             // the head operand is the raw tile: only the tail is computed
    long long tile = blockIdx.x;
    int kb = 0;
    unsigned cnt = 0;
    while (tile < ntiles) {
      const unsigned stage = cnt % TW_STAGES;
      mbar_wait(&tma_full[stage], (cnt / TW_STAGES) & 1);
      // (the tail buffer is free: the TMA warp saw mma_done of this stage's previous use before it started this copy)
      uint8_t* sh = ring + stage * 2 * TW_BLK_BYTES;
      uint8_t* sl = sh + TW_BLK_BYTES;
      if (raw) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t off = tw_sw128(prow + 32 * q, c4);
          const float4 v = *reinterpret_cast<const float4*>(sh + off);
          *reinterpret_cast<float4*>(sl + off) =
              make_float4(tw_tail_trunc(v.x), tw_tail_trunc(v.y), tw_tail_trunc(v.z), tw_tail_trunc(v.w));
        }
      } else {
        const int g = (tile >= tpg) ? 1 : 0;
        const long long row0 = (tile - (long long)g * tpg) * TW_BM;
        const int rows = (int)((a.R - row0 < TW_BM) ? (a.R - row0) : TW_BM);
        const int col = kb * TW_KB + c4 * 4;
        float4 pa4 = make_float4(1.f, 1.f, 1.f, 1.f), pc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.pro) {
          pa4 = *reinterpret_cast<const float4*>(&s_pa[g * 128 + col]);
          pc4 = *reinterpret_cast<const float4*>(&s_pc[g * 128 + col]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int row = prow + 32 * q;
          const uint32_t off = tw_sw128(row, c4);
          const float4 v = *reinterpret_cast<const float4*>(sh + off);
          float t[4] = {v.x, v.y, v.z, v.w};
          if (a.pro) {
            t[0] = fmaf(pa4.x, t[0], pc4.x); t[1] = fmaf(pa4.y, t[1], pc4.y);
            t[2] = fmaf(pa4.z, t[2], pc4.z); t[3] = fmaf(pa4.w, t[3], pc4.w);
            if (a.pro == 2) {
              t[0] = fmaxf(t[0], 0.f); t[1] = fmaxf(t[1], 0.f); t[2] = fmaxf(t[2], 0.f); t[3] = fmaxf(t[3], 0.f);
            }
            if (!(row < rows)) t[0] = t[1] = t[2] = t[3] = 0.f;   // zero-filled rows must not pick up the shift
          }
          float4 h, l;
          tw_split(t[0], h.x, l.x);
          tw_split(t[1], h.y, l.y);
          tw_split(t[2], h.z, l.z);
          tw_split(t[3], h.w, l.w);
          *reinterpret_cast<float4*>(sh + off) = h;
          *reinterpret_cast<float4*>(sl + off) = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);
      if (++kb == nkb) { kb = 0; tile += gridDim.x; }
      ++cnt;
    }
  } else if (warp == 17) {
    // ================================================================================================ TMA issuer
    // one thread runs up to TW_STAGES K-blocks ahead of the MMAs: as soon as the MMAs that read a ring stage have
    // retired it arms the stage's barrier and starts the tensor copy of the next raw K-block into it
    if (lane == 0) {
      unsigned cnt = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int g = (tile >= tpg) ? 1 : 0;
        const long long row0 = (tile - (long long)g * tpg) * TW_BM;
        {   // bulk L2 prefetch of the tile TW_L2_AHEAD rounds ahead (its rows are contiguous)
          const long long pt = tile + (long long)a.l2_ahead * gridDim.x;
          if (a.l2_ahead > 0 && pt < ntiles) {
            const int pg = (pt >= tpg) ? 1 : 0;
            const long long prow0 = (pt - (long long)pg * tpg) * TW_BM;
            const int prows = (int)((a.R - prow0 < TW_BM) ? (a.R - prow0) : TW_BM);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.x + ((long long)pg * a.R + prow0) * a.ldx),
                         "r"((uint32_t)(prows * a.ldx * 4)) : "memory");
          }
        }
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const unsigned stage = cnt % TW_STAGES;
          if (cnt >= TW_STAGES) mbar_wait(&mma_done[stage], ((cnt / TW_STAGES) - 1) & 1);
          mbar_arrive_expect_tx(&tma_full[stage], TW_BLK_BYTES);
          tw_tma_load(ring + stage * 2 * TW_BLK_BYTES, &tmx, kb * TW_KB, (int)row0, g, &tma_full[stage]);
        }
      }
    }
  } else {
    // ================================================================================================= epilogue
    // The accumulator is transposed: TMEM lane = output channel, column = tile row.  Warp e owns channels
    // [32 q, 32 q + 32) (q = e & 3, its TMEM lane quarter) and the tile rows [64 (e >> 2), 64 (e >> 2) + 64); a thread
    // owns ONE channel: bias / ReLU / BatchNorm sums are per-thread scalars, and for each tile row the warp's 32 lanes
    // store 128 contiguous bytes.
    const int e = warp - TW_PROD / 32, q = e & 3;
    const int n = q * 32 + lane;
    const bool chan = n < N;
    const float bias = (a.bias && chan) ? __ldg(a.bias + n) : 0.f;
    double st_s[TW_MAXG] = {0.0, 0.0}, st_q[TW_MAXG] = {0.0, 0.0};
    unsigned ti = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int g = (tile >= tpg) ? 1 : 0;
      const long long row0 = (tile - (long long)g * tpg) * TW_BM;
      const int rows = (int)((a.R - row0 < TW_BM) ? (a.R - row0) : TW_BM);
      const long long base = (long long)g * a.R + row0;
      const uint32_t buf = ti & 1u;
      mbar_wait(&acc_done[buf], (ti >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r0 = (e >> 2) * 64 + hh * 32;          // first tile row of this [32 channels x 32 rows] block
        uint32_t v[32];
        const bool have = (q * 32 < N) && (r0 < rows);   // warp-uniform
        if (have) {
          TW_LD32(v, tmem + buf * 128u + ((uint32_t)(q * 32) << 16) + (uint32_t)r0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (hh == 1) {   // both blocks are in registers / skipped: hand the accumulator back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_free[buf]);
        }
        if (!have) continue;
        float* yp = a.y + (base + r0) * a.ldy + n;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = __uint_as_float(v[j]) + bias;
          if (a.relu) t = fmaxf(t, 0.f);
          if (r0 + j < rows) {                           // warp-uniform: rows past the end of the group are not stored
            if (chan) yp[(long long)j * a.ldy] = t;
            s1 += t;
            s2 = fmaf(t, t, s2);
          }
        }
        if (g == 0) { st_s[0] += (double)s1; st_q[0] += (double)s2; }
        else        { st_s[1] += (double)s1; st_q[1] += (double)s2; }
      }
    }
    if (a.stats && chan) {   // two warps (row halves) per channel, once per kernel
#pragma unroll
      for (int g = 0; g < TW_MAXG; ++g) {
        if (g < a.G) {
          atomicAdd(a.stats + (long long)(g * 2 + 0) * N + n, st_s[g]);
          atomicAdd(a.stats + (long long)(g * 2 + 1) * N + n, st_q[g]);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ------------------
typedef CUresult (*tw_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// tuning knob of the opt-in kernels: SB_TMA_L2PROMO = 0 (none, default) | 1 (64 B) | 2 (128 B) | 3 (256 B)
static CUtensorMapL2promotion tw_l2promo() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SB_TMA_L2PROMO");
    v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
         : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
}
static tw_encode_fn tw_encoder() {
  static tw_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (tw_encode_fn)p;
  }
  return fn;
}
// [G][R][C] fp32 view with row stride ld (floats): dims innermost-first {C, R, G}; box {32, box_rows, 1}; SWIZZLE_128B
static int tw_make_map(CUtensorMap* tm, const float* base, int64_t ld, int64_t R, int32_t G, int32_t C, int box_rows) {
  tw_encode_fn enc = tw_encoder();
  if (!enc) return SB_ERR_UNSUPPORTED;
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)R, (cuuint64_t)G};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)R * (cuuint64_t)ld * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, tw_l2promo(),
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (r == CUDA_SUCCESS) ? SB_OK : SB_ERR_UNSUPPORTED;
}

// Returns SB_ERR_UNSUPPORTED (without setting an error) for anything but the fast shapes; the caller then uses
// linear_tc_kernel (same contract).
int sb_linear_tc_ws_launch(const float* x, int64_t ldx, const float* w, int64_t w_rs, int64_t w_cs, const float* bias,
                            float* y, int64_t ldy, int64_t R, int32_t G, int32_t K, int32_t N, int32_t pro,
                            const float* pa, const float* pc, int32_t relu, double* stats, int32_t accumulate,
                            int32_t ycols, int32_t rawhead, cudaStream_t st) {
  const bool xvec = (ldx % 4 == 0) && ((uintptr_t)x % 16 == 0);
  const bool yvec = (ldy % 4 == 0) && ((uintptr_t)y % 16 == 0);
  if (K < 32 || K > 128 || N < 32 || N > 128 || (K % 32) || (N % 32) || G > TW_MAXG || accumulate || !xvec || !yvec ||
      ycols != N || R * G < 4096 || R >= (1ll << 31))
    return SB_ERR_UNSUPPORTED;
  CUtensorMap tmx;
  if (tw_make_map(&tmx, x, ldx, R, G, K, TW_BM) != SB_OK) return SB_ERR_UNSUPPORTED;
  TwArgs a;
  a.x = x; a.ldx = ldx; a.w = w; a.w_rs = w_rs; a.w_cs = w_cs; a.bias = bias; a.y = y; a.ldy = ldy; a.R = R; a.G = G;
  a.K = K; a.N = N; a.nkb = K / TW_KB;
  a.pro = pro; a.pa = pa; a.pc = pc; a.relu = relu; a.stats = stats; a.rawhead = rawhead;
  {
    static int ahead = -1;
    if (ahead < 0) {
      const char* e = getenv("SB_TMA_L2_AHEAD");
      ahead = (e && e[0] >= '0' && e[0] <= '9') ? atoi(e) : TW_L2_AHEAD;
    }
    a.l2_ahead = ahead;
  }
  const size_t smem = (size_t)TW_STAGES * 2 * TW_BLK_BYTES;
  static bool configured = false;
  if (!configured) {
    SB_CUDA(cudaFuncSetAttribute(linear_tc_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long ntiles = sb_ceil_div(R, TW_BM) * G;
  long long grid = sb_num_sms();
  if (grid > ntiles) grid = ntiles;
  linear_tc_ws_kernel<<<(unsigned)grid, TW_THREADS, smem, st>>>(a, tmx);
  SB_CHECK_LAUNCH("sb_linear_fwd(tcgen05, weight in tensor memory)");
  return SB_OK;
}

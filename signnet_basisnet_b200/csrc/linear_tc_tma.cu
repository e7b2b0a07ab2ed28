// K2 on tcgen05 with the TMA engine on both sides of the tile (third generation of linear_tc.cu, fast shapes only).
//
// Why: the ncu capture of linear_tc_kernel<FAST> (profiles/r1z_linear_tc_ncu.csv + the raw report) shows the LSU data
// pipe of the L1 at 74 % of its peak (411 k wavefronts per SM per launch: shared loads 172 k, shared stores 114 k,
// global 109 k) while DRAM sits at 46 % and the tensor pipe at 36 %: every byte goes global -> registers -> shared on the
// way in (LDG + 2 STS) and TMEM -> registers -> shared -> registers -> global on the way out (STS + LDS + STG), and all of
// these are LSU wavefronts.  This kernel removes the LSU from the two global legs:
//   * input: ONE thread issues a 3-D tensor-map TMA load of the raw fp32 K-block [128 rows x 32 floats] straight into
//     the operand ring in the canonical SWIZZLE_128B layout (rows past the end of a group are zero-filled by the TMA
//     unit, so ragged tiles need no predicates).  The producer warps then only compute the 3xTF32 tail in place
//     (LDS + STS of one buffer); with `rawhead` and no prologue the head operand IS the raw tile (the tensor core reads
//     the top 19 bits of an fp32 word = truncation to tf32; scripts/tc_probe.cu, mode 2, checks exactly this).
//   * output: the epilogue warps write their [32 x 32] block once into a swizzled staging tile and ONE lane issues a
//     tensor-map TMA store (rows past the end of the group are clipped by the TMA unit); the BatchNorm column
//     statistics are still read from the staging tile.
// Arithmetic and MMA order are those of linear_tc_kernel, so with the head rewritten (mode 3) the results are
// bit-identical to it; with the raw head (mode 4) heads are truncated instead of rounded (same error bound).
//
// STATUS: opt-in (sb_set_tensor_cores(3 | 4) or SB_LINEAR_TMA=1 | 2); DESIGN.md says what has been measured.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define TT_BM 128
#define TT_KB 32
#define TT_BLK_BYTES (128 * 128)
#define TT_WORKERS 512
#define TT_PROD 256
#define TT_THREADS (TT_WORKERS + 64)   // + MMA warp (16) + TMA-issue warp (17)
#define TT_MAXG 2
#define TT_ESTAGE_BYTES (8 * 32 * 32 * 4)
#define TT_L2_AHEAD 2

struct TtArgs {
  const float* x;
  long long ldx;
  const float* w;
  long long w_rs, w_cs;
  const float* bias;
  long long R;
  int G, K, N, nkb;
  int pro;
  const float* pa;
  const float* pc;
  int relu;
  double* stats;
  int rawhead;
  int l2_ahead;   // tiles requested into L2 ahead of the TMA loads (SB_TMA_L2_AHEAD, default TT_L2_AHEAD)
};

__device__ __forceinline__ uint64_t tt_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t tt_sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void tt_split(float x, float& h, float& l) {   // == tc_split (linear_tc.cu)
  h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  l = x - h;
}
__device__ __forceinline__ float tt_tail_trunc(float x) {   // x minus the tf32 the tensor core reads from the raw word
  return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}
__device__ __forceinline__ void tt_mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 3-D tensor-map copies (SASS: UTMALDG / UTMASTG); coordinates = (column, row inside the group, group)
__device__ __forceinline__ void tt_tma_load(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tt_tma_store(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
#define TT_LD32(v, taddr)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26," \
               "%27,%28,%29,%30,%31}, [%32];"                                                                         \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),           \
                 "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),           \
                 "=r"(v[30]), "=r"(v[31])                                                                            \
               : "r"(taddr))

__global__ void __launch_bounds__(TT_THREADS, 1)
linear_tc_tma_kernel(const TtArgs a, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = a.nkb;
  uint8_t* Wh = smem;                              // [nkb][16 KB]
  uint8_t* Wl = Wh + nkb * TT_BLK_BYTES;
  uint8_t* ring = Wl + nkb * TT_BLK_BYTES;         // 2 stages x (head 16 KB | tail 16 KB)
  float* estage = reinterpret_cast<float*>(ring + 4 * TT_BLK_BYTES);   // [8 warps][32 rows][32 cols], 4 KB aligned each
  __shared__ __align__(16) float s_pa[TT_MAXG * 128], s_pc[TT_MAXG * 128], s_bias[128];
  __shared__ uint64_t tma_full[2], full[2], mma_done[2], acc_done[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, N = a.N;

  for (int idx = tid; idx < nkb * 128 * TT_KB; idx += TT_THREADS) {
    const int n = idx / (nkb * TT_KB), k = idx - n * (nkb * TT_KB);
    float v = 0.f;
    if (k < K && n < N) v = __ldg(a.w + (long long)n * a.w_rs + (long long)k * a.w_cs);
    float h, l;
    tt_split(v, h, l);
    const uint32_t off = (uint32_t)(k / TT_KB) * TT_BLK_BYTES + tt_sw128(n, (k % TT_KB) >> 2) + (uint32_t)(k & 3) * 4;
    *reinterpret_cast<float*>(Wh + off) = h;
    *reinterpret_cast<float*>(Wl + off) = l;
  }
  for (int idx = tid; idx < TT_MAXG * 128; idx += TT_THREADS) {
    const int g = idx >> 7, c = idx & 127;
    const bool ok = a.pro && g < a.G && c < K;
    s_pa[idx] = ok ? __ldg(a.pa + (long long)g * K + c) : 1.f;
    s_pc[idx] = ok ? __ldg(a.pc + (long long)g * K + c) : 0.f;
  }
  for (int idx = tid; idx < 128; idx += TT_THREADS) s_bias[idx] = (a.bias && idx < N) ? __ldg(a.bias + idx) : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tma_full[i], 1);                   // the issuing thread's arrive.expect_tx + the TMA's bytes
      mbar_init(&full[i], TT_PROD / 32);
      mbar_init(&mma_done[i], 1);
      mbar_init(&acc_done[i], 1);
      mbar_init(&acc_free[i], (TT_WORKERS - TT_PROD) / 32);
    }
    mbar_fence_init();
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const long long tpg = (a.R + TT_BM - 1) / TT_BM;
  const long long ntiles = tpg * a.G;

  if (warp == 16) {
    // =============================================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TT_BM >> 4) << 24);
      unsigned cnt = 0, ti = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        const uint32_t buf = ti & 1u;
        if (ti >= 2) {
          mbar_wait(&acc_free[buf], ((ti >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tacc = tmem + buf * 128u;
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int stage = cnt & 1;
          mbar_wait(&full[stage], (cnt >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = smem_u32(ring + stage * 2 * TT_BLK_BYTES), al = ah + TT_BLK_BYTES;
          const uint32_t wh = smem_u32(Wh + kb * TT_BLK_BYTES), wl = smem_u32(Wl + kb * TT_BLK_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t o = j * 32;
            tt_mma(tacc, tt_make_desc(ah + o), tt_make_desc(wh + o), idesc, (kb | j) ? 1u : 0u);
            tt_mma(tacc, tt_make_desc(ah + o), tt_make_desc(wl + o), idesc, 1u);
            tt_mma(tacc, tt_make_desc(al + o), tt_make_desc(wh + o), idesc, 1u);
          }
          tt_commit(&mma_done[stage]);
          if (kb == nkb - 1) tt_commit(&acc_done[buf]);
        }
      }
    }
  } else if (warp < TT_PROD / 32) {
    // ================================================================================================ producers
    // thread -> 16-byte chunk c4 of rows (tid >> 3) + 32 q, q < 4, of the K-block the TMA unit has put into the ring
    const int prow = tid >> 3, c4 = tid & 7;
    const bool raw = a.rawhead && !a.pro;             // the head operand is the raw tile: only the tail is computed
    struct Cur { long long tile; int kb; };
    auto advance = [&](Cur& c) {
      if (++c.kb == nkb) { c.kb = 0; c.tile += gridDim.x; }
    };
    Cur cur{(long long)blockIdx.x, 0};
    unsigned cnt = 0;
    while (cur.tile < ntiles) {
      const int stage = cnt & 1;
      mbar_wait(&tma_full[stage], (cnt >> 1) & 1);
      // (the tail buffer is free: the TMA warp saw mma_done of this stage's previous use before it started this copy)
      {
        const int g = (cur.tile >= tpg) ? 1 : 0;
        const long long row0 = (cur.tile - (long long)g * tpg) * TT_BM;
        const int rows = (int)((a.R - row0 < TT_BM) ? (a.R - row0) : TT_BM);
        uint8_t* sh = ring + stage * 2 * TT_BLK_BYTES;
        uint8_t* sl = sh + TT_BLK_BYTES;
        if (raw) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t off = tt_sw128(prow + 32 * q, c4);
            const float4 v = *reinterpret_cast<const float4*>(sh + off);
            *reinterpret_cast<float4*>(sl + off) =
                make_float4(tt_tail_trunc(v.x), tt_tail_trunc(v.y), tt_tail_trunc(v.z), tt_tail_trunc(v.w));
          }
        } else {
          const int col = cur.kb * TT_KB + c4 * 4;
          float4 pa4 = make_float4(1.f, 1.f, 1.f, 1.f), pc4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.pro) {
            pa4 = *reinterpret_cast<const float4*>(&s_pa[g * 128 + col]);
            pc4 = *reinterpret_cast<const float4*>(&s_pc[g * 128 + col]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int row = prow + 32 * q;
            const uint32_t off = tt_sw128(row, c4);
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
            tt_split(t[0], h.x, l.x);
            tt_split(t[1], h.y, l.y);
            tt_split(t[2], h.z, l.z);
            tt_split(t[3], h.w, l.w);
            *reinterpret_cast<float4*>(sh + off) = h;
            *reinterpret_cast<float4*>(sl + off) = l;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);
      advance(cur);
      ++cnt;
    }
  } else if (warp == 17) {
    // ================================================================================================ TMA issuer
    // one thread runs up to two K-blocks ahead of the MMAs: as soon as the MMAs that read a ring stage have retired it
    // arms the stage's barrier and starts the tensor copy of the next raw K-block into it
    if (lane == 0) {
      unsigned cnt = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int g = (tile >= tpg) ? 1 : 0;
        const long long row0 = (tile - (long long)g * tpg) * TT_BM;
        {   // bulk L2 prefetch of the tile TT_L2_AHEAD rounds ahead (its rows are contiguous)
          const long long pt = tile + (long long)a.l2_ahead * gridDim.x;
          if (a.l2_ahead > 0 && pt < ntiles) {
            const int pg = (pt >= tpg) ? 1 : 0;
            const long long prow0 = (pt - (long long)pg * tpg) * TT_BM;
            const int prows = (int)((a.R - prow0 < TT_BM) ? (a.R - prow0) : TT_BM);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.x + ((long long)pg * a.R + prow0) * a.ldx),
                         "r"((uint32_t)(prows * a.ldx * 4)) : "memory");
          }
        }
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int stage = cnt & 1;
          if (cnt >= 2) mbar_wait(&mma_done[stage], ((cnt >> 1) - 1) & 1);
          mbar_arrive_expect_tx(&tma_full[stage], TT_BLK_BYTES);
          tt_tma_load(ring + stage * 2 * TT_BLK_BYTES, &tmx, kb * TT_KB, (int)row0, g, &tma_full[stage]);
        }
      }
    }
  } else {
    // ================================================================================================= epilogue
    const int e = warp - TT_PROD / 32, eq = e & 3;
    float* wst = estage + e * (32 * 32);               // [32 rows][32 cols], 16-byte chunks XOR-swizzled by (row & 7)
    double st_s[2][TT_MAXG] = {{0.0, 0.0}, {0.0, 0.0}}, st_q[2][TT_MAXG] = {{0.0, 0.0}, {0.0, 0.0}};
    unsigned ti = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int g = (tile >= tpg) ? 1 : 0;
      const long long row0 = (tile - (long long)g * tpg) * TT_BM;
      const int rows = (int)((a.R - row0 < TT_BM) ? (a.R - row0) : TT_BM);
      const uint32_t buf = ti & 1u;
      mbar_wait(&acc_done[buf], (ti >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = eq * 32 + lane;
      const bool live = row < rows;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int ec0 = ((e >> 2) + 2 * hh) * 32;
        uint32_t v[32];
        if (ec0 < N) {
          TT_LD32(v, tmem + buf * 128u + ((uint32_t)(eq * 32) << 16) + (uint32_t)ec0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (hh == 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_free[buf]);
        }
        if (ec0 >= N) continue;
        // the TMA store of the previous block has finished READING the staging tile
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[ec0 + i * 4]);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float t = __uint_as_float(v[i * 4 + j]) + bb[j];
            if (a.relu) t = fmaxf(t, 0.f);
            o[j] = live ? t : 0.f;
          }
          *reinterpret_cast<float4*>(wst + lane * 32 + ((i ^ (lane & 7)) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA unit
        __syncwarp();
        if (lane == 0 && eq * 32 < rows) {
          tt_tma_store(&tmy, wst, ec0, (int)row0 + eq * 32, g);       // rows >= R of the group are clipped
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (a.stats) {   // lane -> column ec0 + lane over the block's 32 rows (conflict-free LDS.32; dead rows hold 0)
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const float t = wst[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))];
            s1 += t;
            s2 = fmaf(t, t, s2);
          }
          if (g == 0) { st_s[hh][0] += (double)s1; st_q[hh][0] += (double)s2; }
          else        { st_s[hh][1] += (double)s1; st_q[hh][1] += (double)s2; }
        }
        __syncwarp();
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA exits
    if (a.stats) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = ((e >> 2) + 2 * hh) * 32 + lane;
        if (col < N) {
#pragma unroll
          for (int g = 0; g < TT_MAXG; ++g) {
            if (g < a.G) {
              atomicAdd(a.stats + (long long)(g * 2 + 0) * N + col, st_s[hh][g]);
              atomicAdd(a.stats + (long long)(g * 2 + 1) * N + col, st_q[hh][g]);
            }
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ------------------
typedef CUresult (*tt_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// tuning knob of the opt-in kernels: SB_TMA_L2PROMO = 0 (none, default) | 1 (64 B) | 2 (128 B) | 3 (256 B)
static CUtensorMapL2promotion tt_l2promo() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SB_TMA_L2PROMO");
    v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
         : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
}
static tt_encode_fn tt_encoder() {
  static tt_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (tt_encode_fn)p;
  }
  return fn;
}
// [G][R][C] fp32 view with row stride ld (floats): dims innermost-first {C, R, G}; box {32, box_rows, 1}; SWIZZLE_128B
static int tt_make_map(CUtensorMap* tm, const float* base, int64_t ld, int64_t R, int32_t G, int32_t C, int box_rows) {
  tt_encode_fn enc = tt_encoder();
  if (!enc) return SB_ERR_UNSUPPORTED;
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)R, (cuuint64_t)G};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)R * (cuuint64_t)ld * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, tt_l2promo(),
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (r == CUDA_SUCCESS) ? SB_OK : SB_ERR_UNSUPPORTED;
}

// Returns SB_ERR_UNSUPPORTED (without setting an error) for anything but the fast shapes; the caller then uses
// linear_tc_kernel (same contract).
int sb_linear_tc_tma_launch(const float* x, int64_t ldx, const float* w, int64_t w_rs, int64_t w_cs, const float* bias,
                            float* y, int64_t ldy, int64_t R, int32_t G, int32_t K, int32_t N, int32_t pro,
                            const float* pa, const float* pc, int32_t relu, double* stats, int32_t accumulate,
                            int32_t ycols, int32_t rawhead, cudaStream_t st) {
  const bool xvec = (ldx % 4 == 0) && ((uintptr_t)x % 16 == 0);
  const bool yvec = (ldy % 4 == 0) && ((uintptr_t)y % 16 == 0);
  if (K < 32 || K > 128 || N < 32 || N > 128 || (K % 32) || (N % 32) || G > TT_MAXG || accumulate || !xvec || !yvec ||
      ycols != N || R * G < 4096 || R >= (1ll << 31))
    return SB_ERR_UNSUPPORTED;
  CUtensorMap tmx, tmy;
  if (tt_make_map(&tmx, x, ldx, R, G, K, TT_BM) != SB_OK || tt_make_map(&tmy, y, ldy, R, G, N, 32) != SB_OK)
    return SB_ERR_UNSUPPORTED;
  TtArgs a;
  a.x = x; a.ldx = ldx; a.w = w; a.w_rs = w_rs; a.w_cs = w_cs; a.bias = bias; a.R = R; a.G = G;
  a.K = K; a.N = N; a.nkb = K / TT_KB;
  a.pro = pro; a.pa = pa; a.pc = pc; a.relu = relu; a.stats = stats; a.rawhead = rawhead;
  {
    static int ahead = -1;
    if (ahead < 0) {
      const char* e = getenv("SB_TMA_L2_AHEAD");
      ahead = (e && e[0] >= '0' && e[0] <= '9') ? atoi(e) : TT_L2_AHEAD;
    }
    a.l2_ahead = ahead;
  }
  const size_t smem = (size_t)2 * a.nkb * TT_BLK_BYTES + 4 * TT_BLK_BYTES + TT_ESTAGE_BYTES;
  static bool configured = false;
  if (!configured) {
    const int mx = 2 * 4 * TT_BLK_BYTES + 4 * TT_BLK_BYTES + TT_ESTAGE_BYTES;
    SB_CUDA(cudaFuncSetAttribute(linear_tc_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    configured = true;
  }
  const long long ntiles = sb_ceil_div(R, TT_BM) * G;
  long long grid = sb_num_sms();
  if (grid > ntiles) grid = ntiles;
  linear_tc_tma_kernel<<<(unsigned)grid, TT_THREADS, smem, st>>>(a, tmx, tmy);
  SB_CHECK_LAUNCH("sb_linear_fwd(tcgen05 + TMA)");
  return SB_OK;
}

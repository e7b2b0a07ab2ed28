// K2 on a CTA PAIR (thread-block cluster of 2, tcgen05 cta_group::2): the same contract and the same arithmetic as
// linear_tc_kernel<FAST> (linear_tc.cu: y = f(x) W^T + bias, 3xTF32 error-compensated, fp32 accumulation in tensor
// memory, BatchNorm column statistics in the epilogue) with one UMMA of M = 256 spanning two SMs:
//   * each CTA supplies ITS OWN 128-row tile of x as the A operand and drains ITS OWN [128 x N] accumulator;
//   * the weight (B operand) is split across the pair: a CTA keeps only N/2 of its rows resident (head + tail =
//     2 x 32 KB at K = N = 128 instead of 2 x 64 KB), which is what pays for a 4-stage operand ring instead of the 2
//     stages the single-CTA kernel has room for (shared memory is exactly full there);
//   * one thread of the leader CTA issues every tcgen05.mma for both SMs; full[] / acc_free[] live in the leader and
//     collect arrivals from both CTAs (mapa + cluster-scope mbarrier.arrive), mma_done[] / acc_done[] are signalled in
//     both CTAs by one multicast tcgen05.commit.
// The barrier protocol is the one validated on the B200 by scripts/tc_probe_2cta_pipe.cu
// (profiles/r1z_tc_probe_2cta_pipe.log).  Shapes outside the fast path (K, N multiples of 32 up to 128, 16-byte
// aligned rows, no accumulate) stay on the single-CTA kernel.
//
// STATUS: opt-in (sb_set_tensor_cores(2) or SB_LINEAR_PAIR=1); see DESIGN.md for what has and has not been measured.
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define TP_BM 128
#define TP_KB 32                        // floats per K-block = one 128-byte swizzle row
#define TP_A_BLK (128 * 128)            // [128 rows x 32 floats] activation block
#define TP_B_BLK (64 * 128)             // [<= 64 weight rows x 32 floats]: this CTA's half of one weight K-block
#define TP_STAGES 4
#define TP_WORKERS 512
#define TP_PROD 256                     // producer threads (warps 0..7); warps 8..15 = epilogue; warp 16 = MMA
#define TP_THREADS (TP_WORKERS + 32)
#define TP_MAXG 2
#define TP_ESTAGE_BYTES (8 * 32 * 32 * 4)
#define TP_L2_AHEAD 2
static_assert((TP_STAGES & (TP_STAGES - 1)) == 0 && TP_STAGES >= 2, "ring depth: power of two (stage = cnt & mask)");

struct TpArgs {
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
};

__device__ __forceinline__ uint64_t tp_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;     // 1024 B between 8-row groups
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t tp_sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void tp_split(float x, float& h, float& l) {   // == tc_split (linear_tc.cu)
  h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  l = x - h;
}
__device__ __forceinline__ void tp_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in the LEADER CTA (cluster rank 0)
__device__ __forceinline__ void tp_arrive_leader(uint64_t* b) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(b)), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool tp_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tp_wait(uint64_t* bar, uint32_t parity) {   // bounded: trap, never hang the GPU
  if (tp_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (!tp_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0 && clock64() - t0 > 4000000000ll) {
      printf("libsignnet_b200: pair-kernel mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tp_mma2(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tp_commit_pair(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
#define TP_LD32(v, taddr)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26," \
               "%27,%28,%29,%30,%31}, [%32];"                                                                         \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),           \
                 "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),           \
                 "=r"(v[30]), "=r"(v[31])                                                                            \
               : "r"(taddr))

// Work decomposition: the 128-row tiles of the [G x R] row space are numbered as in linear_tc_kernel (group-major, no
// tile straddles a group); pair p covers tiles 2p (leader) and 2p + 1 (peer).  With an odd tile count the peer's last
// tile is dead: it feeds zeros, skips its stores and still takes part in every barrier.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TP_THREADS, 1) linear_tc_pair_kernel(const TpArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = a.nkb;
  uint8_t* Wh = smem;                                  // [nkb][8 KB]: rows rank * N/2 .. + N/2 of the weight, heads
  uint8_t* Wl = Wh + nkb * TP_B_BLK;                   // tails
  uint8_t* ring = Wl + nkb * TP_B_BLK;                 // TP_STAGES x (head 16 KB | tail 16 KB)
  float* estage = reinterpret_cast<float*>(ring + TP_STAGES * 2 * TP_A_BLK);
  __shared__ __align__(16) float s_pa[TP_MAXG * 128], s_pc[TP_MAXG * 128], s_bias[128];
  __shared__ uint64_t full[TP_STAGES], mma_done[TP_STAGES], acc_done[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;

  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, N = a.N, NH = a.N >> 1;
  const long long cluster = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  // ---- one-time: this CTA's half of the weight, split into head/tail, canonical K-major SW128 blocks
  for (int idx = tid; idx < nkb * 64 * TP_KB; idx += TP_THREADS) {
    const int n = idx / (nkb * TP_KB), k = idx - n * (nkb * TP_KB);   // lanes along k: conflict-free swizzle rows
    float v = 0.f;
    if (k < K && n < NH) v = __ldg(a.w + (long long)(rank * NH + n) * a.w_rs + (long long)k * a.w_cs);
    float h, l;
    tp_split(v, h, l);
    const uint32_t off = (uint32_t)(k / TP_KB) * TP_B_BLK + tp_sw128(n, (k % TP_KB) >> 2) + (uint32_t)(k & 3) * 4;
    *reinterpret_cast<float*>(Wh + off) = h;
    *reinterpret_cast<float*>(Wl + off) = l;
  }
  for (int idx = tid; idx < TP_MAXG * 128; idx += TP_THREADS) {
    const int g = idx >> 7, c = idx & 127;
    const bool ok = a.pro && g < a.G && c < K;
    s_pa[idx] = ok ? __ldg(a.pa + (long long)g * K + c) : 1.f;
    s_pc[idx] = ok ? __ldg(a.pc + (long long)g * K + c) : 0.f;
  }
  for (int idx = tid; idx < 128; idx += TP_THREADS) s_bias[idx] = (a.bias && idx < N) ? __ldg(a.bias + idx) : 0.f;
  if (tid == 0) {
    for (int i = 0; i < TP_STAGES; ++i) {
      mbar_init(&full[i], 2 * (TP_PROD / 32));        // producer warps of BOTH CTAs (only the leader's copy is used)
      mbar_init(&mma_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_done[i], 1);
      mbar_init(&acc_free[i], 2 * ((TP_WORKERS - TP_PROD) / 32));   // epilogue warps of both CTAs (leader's copy)
    }
    mbar_fence_init();
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  tp_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const long long tpg = (a.R + TP_BM - 1) / TP_BM;
  const long long ntiles = tpg * a.G;
  const long long npairs = (ntiles + 1) >> 1;

  if (warp == 16) {
    // ============================================================================== MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      const uint32_t idesc =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      unsigned cnt = 0, ti = 0;
      for (long long p = cluster; p < npairs; p += nclusters, ++ti) {
        const uint32_t buf = ti & 1u;
        if (ti >= 2) {   // the epilogue warps of both CTAs have drained the accumulator this tile reuses
          tp_wait(&acc_free[buf], ((ti >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tacc = tmem + buf * 128u;
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int stage = cnt & (TP_STAGES - 1);
          tp_wait(&full[stage], (cnt / TP_STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = smem_u32(ring + stage * 2 * TP_A_BLK), al = ah + TP_A_BLK;
          const uint32_t wh = smem_u32(Wh + kb * TP_B_BLK), wl = smem_u32(Wl + kb * TP_B_BLK);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t o = j * 32;
            tp_mma2(tacc, tp_make_desc(ah + o), tp_make_desc(wh + o), idesc, (kb | j) ? 1u : 0u);
            tp_mma2(tacc, tp_make_desc(ah + o), tp_make_desc(wl + o), idesc, 1u);
            tp_mma2(tacc, tp_make_desc(al + o), tp_make_desc(wh + o), idesc, 1u);
          }
          tp_commit_pair(&mma_done[stage]);
          if (kb == nkb - 1) tp_commit_pair(&acc_done[buf]);
        }
      }
    }
  } else if (warp < TP_PROD / 32) {
    // ================================================================================================ producers
    // thread -> float4 column c4 of rows (tid >> 3) + 32 q, q < 4, of every K-block of this CTA's tile
    const int prow = tid >> 3, c4 = tid & 7;
    const float* xthread = a.x + (long long)prow * a.ldx + c4 * 4;
    const long long xstride32 = 32 * a.ldx;
    float4 pre[2][4];                                  // two K-blocks in flight (slot = position parity)
    struct Cur { long long p; int kb; };
    auto advance = [&](Cur& c) {
      if (++c.kb == nkb) { c.kb = 0; c.p += nclusters; }
    };
    // rows of this CTA's tile of pair p (0 for the dead tile of an odd tile count), its group and first row
    auto tile_of = [&](long long p, int& g, long long& base) -> int {
      const long long tile = 2 * p + rank;
      g = 0; base = 0;
      if (tile >= ntiles) return 0;
      g = (tile >= tpg) ? 1 : 0;
      const long long row0 = (tile - (long long)g * tpg) * TP_BM;
      base = (long long)g * a.R + row0;
      return (int)((a.R - row0 < TP_BM) ? (a.R - row0) : TP_BM);
    };
    auto load_block = [&](const Cur& c, float4 (&dst)[4]) {
      if (c.p < npairs) {
        int g; long long base;
        const int rows = tile_of(c.p, g, base);
        const float* pb = xthread + base * a.ldx + c.kb * TP_KB;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = (prow + 32 * q < rows) ? ldg4(pb + q * xstride32) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_block = [&](int stage, const Cur& c, const float4 (&pv)[4]) {
      int g; long long base;
      const int rows = tile_of(c.p, g, base);
      uint8_t* sh = ring + stage * 2 * TP_A_BLK;
      uint8_t* sl = sh + TP_A_BLK;
      const int col = c.kb * TP_KB + c4 * 4;
      float4 pa4 = make_float4(1.f, 1.f, 1.f, 1.f), pc4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.pro) {
        pa4 = *reinterpret_cast<const float4*>(&s_pa[g * 128 + col]);
        pc4 = *reinterpret_cast<const float4*>(&s_pc[g * 128 + col]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int row = prow + 32 * q;
        float t[4] = {pv[q].x, pv[q].y, pv[q].z, pv[q].w};
        if (a.pro) {
          t[0] = fmaf(pa4.x, t[0], pc4.x); t[1] = fmaf(pa4.y, t[1], pc4.y);
          t[2] = fmaf(pa4.z, t[2], pc4.z); t[3] = fmaf(pa4.w, t[3], pc4.w);
          if (a.pro == 2) {
            t[0] = fmaxf(t[0], 0.f); t[1] = fmaxf(t[1], 0.f); t[2] = fmaxf(t[2], 0.f); t[3] = fmaxf(t[3], 0.f);
          }
        }
        if (!(row < rows)) t[0] = t[1] = t[2] = t[3] = 0.f;   // dead rows must not pick up the prologue's shift
        float4 h, l;
        tp_split(t[0], h.x, l.x);
        tp_split(t[1], h.y, l.y);
        tp_split(t[2], h.z, l.z);
        tp_split(t[3], h.w, l.w);
        const uint32_t off = tp_sw128(row, c4);
        *reinterpret_cast<float4*>(sh + off) = h;
        *reinterpret_cast<float4*>(sl + off) = l;
      }
    };
    auto l2_ahead = [&](long long p) {   // one bulk prefetch per tile: its rows are contiguous
      const long long pp = p + (long long)TP_L2_AHEAD * nclusters;
      if (pp < npairs) {
        int g; long long base;
        const int rows = tile_of(pp, g, base);
        if (rows > 0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.x + base * a.ldx),
                       "r"((uint32_t)(rows * a.ldx * 4)) : "memory");
      }
    };
    auto publish = [&](int stage) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tp_arrive_leader(&full[stage]);
    };

    Cur cur{cluster, 0}, nxt = cur;
    load_block(nxt, pre[0]);
    advance(nxt);
    load_block(nxt, pre[1]);
    advance(nxt);
    unsigned cnt = 0;
#define TP_POSITION(SLOT)                                                                             \
    {                                                                                                 \
      const int stage = cnt & (TP_STAGES - 1);                                                        \
      if (cnt >= TP_STAGES) tp_wait(&mma_done[stage], ((cnt / TP_STAGES) - 1) & 1);                   \
      if (tid == 0 && cur.kb == 0) l2_ahead(cur.p);                                                   \
      store_block(stage, cur, pre[SLOT]);                                                             \
      publish(stage);                                                                                 \
      load_block(nxt, pre[SLOT]);                                                                     \
      advance(nxt);                                                                                   \
      advance(cur);                                                                                   \
      ++cnt;                                                                                          \
    }
    while (cur.p < npairs) {
      TP_POSITION(0)
      if (cur.p >= npairs) break;
      TP_POSITION(1)
    }
#undef TP_POSITION
  } else {
    // ================================================================================================= epilogue
    // warp e owns TMEM lanes [32 q, 32 q + 32) of THIS CTA's accumulator and the column blocks (e >> 2), (e >> 2) + 2
    const int e = warp - TP_PROD / 32, eq = e & 3;
    float* wst = estage + e * (32 * 32);               // [32 rows][32 cols], 16-byte chunks XOR-swizzled by (row & 7)
    const long long ystride4 = 4 * a.ldy;
    double st_s[2][TP_MAXG] = {{0.0, 0.0}, {0.0, 0.0}}, st_q[2][TP_MAXG] = {{0.0, 0.0}, {0.0, 0.0}};
    unsigned ti = 0;
    for (long long p = cluster; p < npairs; p += nclusters, ++ti) {
      const long long tile = 2 * p + rank;
      const bool dead = tile >= ntiles;
      const int g = (!dead && tile >= tpg) ? 1 : 0;
      const long long row0 = dead ? 0 : (tile - (long long)g * tpg) * TP_BM;
      const int rows = dead ? 0 : (int)((a.R - row0 < TP_BM) ? (a.R - row0) : TP_BM);
      const long long base = (long long)g * a.R + row0;
      const uint32_t buf = ti & 1u;
      tp_wait(&acc_done[buf], (ti >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = eq * 32 + lane;
      const bool live = row < rows;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int ec0 = ((e >> 2) + 2 * hh) * 32;
        const bool have = (ec0 < N) && !dead;
        uint32_t v[32];
        if (have) {
          TP_LD32(v, tmem + buf * 128u + ((uint32_t)(eq * 32) << 16) + (uint32_t)ec0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (hh == 1) {   // both blocks are in registers: hand the accumulator back to the leader's MMA thread
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) tp_arrive_leader(&acc_free[buf]);
        }
        if (!have) continue;
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
        __syncwarp();
        {   // coalesced write-back: 4 rows x 128 contiguous bytes per instruction
          const int c = lane & 7;
          float* ybase = a.y + (base + eq * 32 + (lane >> 3)) * a.ldy + ec0 + c * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + (lane >> 3);
            if (eq * 32 + r < rows)
              *reinterpret_cast<float4*>(ybase + it * ystride4) =
                  *reinterpret_cast<const float4*>(wst + r * 32 + ((c ^ (r & 7)) << 2));
          }
        }
        if (a.stats) {   // lane -> column ec0 + lane over the block's 32 rows (conflict-free LDS.32)
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
    if (a.stats) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = ((e >> 2) + 2 * hh) * 32 + lane;
        if (col < N) {
#pragma unroll
          for (int g = 0; g < TP_MAXG; ++g) {
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
  tp_cluster_sync();   // the leader's MMAs read the peer's shared memory: neither CTA may leave before both are done
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// Returns SB_ERR_UNSUPPORTED (without setting an error) for anything but the fast shapes; the caller then uses the
// single-CTA kernel (same results, same contract).
int sb_linear_tc_pair_launch(const float* x, int64_t ldx, const float* w, int64_t w_rs, int64_t w_cs,
                             const float* bias, float* y, int64_t ldy, int64_t R, int32_t G, int32_t K, int32_t N,
                             int32_t pro, const float* pa, const float* pc, int32_t relu, double* stats,
                             int32_t accumulate, int32_t ycols, cudaStream_t st) {
  const bool xvec = (ldx % 4 == 0) && ((uintptr_t)x % 16 == 0);
  const bool yvec = (ldy % 4 == 0) && ((uintptr_t)y % 16 == 0);
  if (K < 32 || K > 128 || N < 32 || N > 128 || (K % 32) || (N % 32) || G > TP_MAXG || accumulate || !xvec || !yvec ||
      ycols != N || R * G < 8192)
    return SB_ERR_UNSUPPORTED;
  const int sms = sb_num_sms();
  if (sms < 2) return SB_ERR_UNSUPPORTED;
  TpArgs a;
  a.x = x; a.ldx = ldx; a.w = w; a.w_rs = w_rs; a.w_cs = w_cs; a.bias = bias; a.y = y; a.ldy = ldy; a.R = R; a.G = G;
  a.K = K; a.N = N; a.nkb = K / TP_KB;
  a.pro = pro; a.pa = pa; a.pc = pc; a.relu = relu; a.stats = stats;
  const size_t smem = (size_t)2 * a.nkb * TP_B_BLK + (size_t)TP_STAGES * 2 * TP_A_BLK + TP_ESTAGE_BYTES;
  static bool configured = false;
  if (!configured) {
    const int mx = 2 * 4 * TP_B_BLK + TP_STAGES * 2 * TP_A_BLK + TP_ESTAGE_BYTES;
    SB_CUDA(cudaFuncSetAttribute(linear_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    configured = true;
  }
  const long long ntiles = sb_ceil_div(R, TP_BM) * G;
  const long long npairs = (ntiles + 1) / 2;
  long long nclusters = sms / 2;
  if (nclusters > npairs) nclusters = npairs;
  linear_tc_pair_kernel<<<(unsigned)(2 * nclusters), TP_THREADS, smem, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_fwd(tcgen05 pair)");
  return SB_OK;
}

// K1 + K2 fused: the GIN neighbourhood aggregate of a phi layer and the first Linear of its MaskedMLP in ONE kernel.
//     A[s, (b,j,i), :] = (1 + eps) X[s, (b,j,i), :] + sum_{(u -> i) in E_b} X[s, (b,j,u), :]         (K1, gin_agg.cu)
//     H = A W0^T   (+ per-(sign, channel) column sums of H and H^2 for the BatchNorm that follows)     (K2, linear_tc.cu)
// Reference: MaskedGINConv.forward = `self.layer(x, edge_index)` then `self.nn(x, mask)`
// (Alchemy/sign_net/model_utils/masked_layers.py:74-84; MaskedMLP first layer :54-58); BASELINE.json's north_star names
// exactly this pair ("a fused CSR scatter-add + MLP kernel ... TMA staging of node-feature tiles into shared memory").
//
// Unfused, K1 writes A (one activation tensor T) and K2 reads it back: 4T of HBM traffic for the pair.  Here the
// aggregated tile never leaves the SM on its way into the contraction: read X (1T), write A once for the backward's
// weight gradient (1T) and H (1T) = 3T, and one kernel launch instead of two.
//
// One persistent CTA per SM, four warp roles connected by mbarriers (no block-wide barrier inside the tile loop):
//   warp 0        producer: unit record + packed neighbour words + cp.async.bulk (TMA) of the X tile - a (sign, graph,
//                 slot-chunk) tile of <= 64 rows x 128 floats, closed under the neighbourhood relation - into a 3-stage
//                 ring (exactly gin_agg.cu's producer);
//   warps 2..9    aggregators: neighbour sums out of shared memory in CSR order (bit-identical to gin_agg.cu), stream
//                 the A row to global memory, split it into tf32 head + tail (3xTF32, as linear_tc.cu) and st.shared
//                 both, 128B-swizzled K-major, into one of two operand buffers;
//   warps 1, 18   MMA issuers (K-blocks 0-1 / 2-3): tcgen05.mma kind::tf32, D[128 channels x N rows] += W(head|tail, resident in TENSOR
//                 MEMORY for the whole kernel, lane = output channel) * A^T (shared memory), N = the tile's row count
//                 rounded up to 16; two 64-column accumulators per issuer rotate, the epilogue adds the two partial sums;
//   warps 10..17  epilogue: tcgen05.ld of [32 channels x 32 rows] blocks; a thread owns one channel, so the BatchNorm
//                 sums are per-thread scalars and every tile row leaves as one coalesced 128-byte store.
//
// MEASURED (B200, cfg 4 size: 2 x 575 454 rows; profiles/r2_fused_agg_linear_ncu.csv, r2_fused_bench.log): A bit-identical
// to gin_agg.cu, H equal to linear_tc.cu's to fp32 rounding; 411-417 us against 207 + 280 = 489 us for the two kernels =
// 0.66 of the HBM peak on its 3T of traffic (0.59 inside the power-capped step).  With ONE issuing thread it was 464-477
// us: a thread issues a tcgen05.mma only every ~100 cycles while the tensor core needs 48 at N = 64
// (scripts/mma_rate2.cu), so the 48 instructions of a <= 64-row tile (4 K-blocks x 4 k-steps x 3 products) cost 5 000
// cycles of issue next to 3 000 of HBM time and the aggregators sat on `opfree`; hence the two issuing warps.  Shared
// memory cannot hold a 256-row 3xTF32 operand (256 KB) next to the tile ring.  (16 aggregator + 4 epilogue warps: 592 us.)
// Shapes: row stride ld = K = 128 floats (the phi stack at n_hid = 128, every layer but the first), h <= 128 output
// channels.  Everything else returns SB_ERR_UNSUPPORTED and the caller runs sb_gin_agg + sb_linear_fwd.
// Tensor memory: accumulators 2 issuers x 2 x 64 columns + weight head 128 + weight tail 128 = 512 columns.
// Shared memory: operand buffers 2 x 64 KB + X ring 3 x 32 KB + neighbour words + zero row = 226 KB.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define GF_LD 128
#define GF_TILE_BYTES 32768
#define GF_ROWS 64
#define GF_XSTAGES 3
#define GF_KB_BYTES (GF_ROWS * 128)       // one K-block operand buffer: [64 rows][32 floats], 8-row swizzle atoms
#define GF_OP_BYTES (8 * GF_KB_BYTES)     // 4 K-blocks of heads, then 4 K-blocks of tails
#define GF_OPBUFS 2
#define GF_ACCS 2                        // accumulators per MMA issuer (tiles in flight between MMA and epilogue)
#define GF_ISSUERS 2
#define GF_AGG_WARPS 8
#define GF_EPI_WARPS 8
#define GF_FIRST_AGG 2
#define GF_FIRST_EPI (GF_FIRST_AGG + GF_AGG_WARPS)
#define GF_ISSUER2 (GF_FIRST_EPI + GF_EPI_WARPS)      // the second MMA-issuing warp
#define GF_THREADS (32 * (GF_ISSUER2 + 1))
#define GF_NB_MAXN 128
#define GF_COL_WH 256
#define GF_COL_WL 384
#define GF_NB_SLOW 0xFEu
#define GF_NB_NONE 0xFFu
#define GF_DYN_SMEM (GF_OPBUFS * GF_OP_BYTES + GF_XSTAGES * GF_TILE_BYTES + GF_XSTAGES * GF_NB_MAXN * 4 + GF_LD * 4)

struct GfArgs {
  const float* x;
  float* a_out;
  float* h_out;
  double* stats;          // [S][2][h] fp64 or null
  const float* eps;       // device scalar or null
  const float* w;
  long long w_rs, w_cs;
  const int32_t* unit_ptr;
  const int32_t* unit_desc;
  const uint32_t* nbr_pack;
  const int32_t* nbr_ptr;
  const int32_t* nbr_idx;
  long long R;
  int B, S, h, ldh;
};

struct GfDesc {
  long long row0;
  int n, rows, node0, nb_local;
  unsigned magic;
  int pad;
};

__device__ __forceinline__ uint64_t gf_make_desc(uint32_t saddr) {   // canonical K-major SWIZZLE_128B (linear_tc.cu)
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t gf_sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void gf_split(float x, float& h, float& l) {   // == tc_split (linear_tc.cu)
  h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  l = x - h;
}
// A operand from tensor memory (lane = output channel, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void gf_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void gf_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#define GF_LD32(v, taddr)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26," \
               "%27,%28,%29,%30,%31}, [%32];"                                                                         \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),       \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),           \
                 "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),           \
                 "=r"(v[30]), "=r"(v[31])                                                                            \
               : "r"(taddr))
#define GF_ST32(taddr, v)                                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                       \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27," \
               "%28,%29,%30,%31,%32};"                                                                                \
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),   \
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),         \
                 "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),       \
                 "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])        \
               : "memory")

__device__ __forceinline__ float4 gf_add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 gf_lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__global__ void __launch_bounds__(GF_THREADS, 1) gin_lin_fused_kernel(const GfArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* opbuf = smem;                                                   // [GF_OPBUFS][GF_OP_BYTES]
  uint8_t* xring = smem + GF_OPBUFS * GF_OP_BYTES;                         // [GF_XSTAGES][GF_TILE_BYTES]
  uint32_t* nbw = reinterpret_cast<uint32_t*>(xring + GF_XSTAGES * GF_TILE_BYTES);   // [GF_XSTAGES][GF_NB_MAXN]
  float* zrow = reinterpret_cast<float*>(nbw + GF_XSTAGES * GF_NB_MAXN);   // GF_LD zeros
  __shared__ uint64_t xfull[GF_XSTAGES], xempty[GF_XSTAGES], opfull[GF_OPBUFS], opfree[GF_OPBUFS], accdone[GF_ACCS],
      accfree[GF_ACCS];
  __shared__ GfDesc descs[GF_XSTAGES];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int U = __ldg(a.unit_ptr + a.B);
  const long long total = (long long)U * a.S;

  for (int i = tid; i < GF_LD; i += GF_THREADS) zrow[i] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < GF_XSTAGES; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], GF_AGG_WARPS); }
    for (int s = 0; s < GF_OPBUFS; ++s) { mbar_init(&opfull[s], GF_AGG_WARPS); mbar_init(&opfree[s], GF_ISSUERS); }
    for (int s = 0; s < GF_ACCS; ++s) { mbar_init(&accdone[s], GF_ISSUERS); mbar_init(&accfree[s], GF_EPI_WARPS); }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  // ---- one-time: the split weight into tensor memory (epilogue warps; lane quarter q = warp & 3; each writes heads and
  // tails of its 32 channels; thread -> output channel n = 32 q + lane; channels >= h hold zeros)
  if (warp >= GF_FIRST_EPI && warp < GF_ISSUER2) {
    const int q = warp & 3, is_tail = ((warp - GF_FIRST_EPI) >> 2) & 1;
    const int n = q * 32 + lane;
    for (int cb = 0; cb < GF_LD / 32; ++cb) {
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int k = cb * 32 + j;
        float wv = 0.f;
        if (n < a.h) wv = __ldg(a.w + (long long)n * a.w_rs + (long long)k * a.w_cs);
        float hh, ll;
        gf_split(wv, hh, ll);
        v[j] = __float_as_uint(is_tail ? ll : hh);
      }
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((is_tail ? GF_COL_WL : GF_COL_WH) + cb * 32);
      GF_ST32(taddr, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 0) {
    // =========================================================================== producer (gin_agg.cu's, T = 1)
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
      const int local = (n <= GF_NB_MAXN) ? 1 : 0;
      uint32_t nv[GF_NB_MAXN / 32];
      if (local) {
#pragma unroll
        for (int q = 0; q < GF_NB_MAXN / 32; ++q) {
          const int i = lane + 32 * q;
          nv[q] = (i < n) ? __ldg(a.nbr_pack + node0 + i) : 0u;
        }
      }
      mbar_wait(&xempty[stage], phase ^ 1u);
      if (local) {
        uint32_t* dst = nbw + stage * GF_NB_MAXN;
#pragma unroll
        for (int q = 0; q < GF_NB_MAXN / 32; ++q) {
          const int i = lane + 32 * q;
          if (i < n) dst[i] = nv[q];
        }
      }
      __syncwarp();
      if (lane == 0) {
        GfDesc d;
        d.row0 = row0; d.n = n; d.rows = rows; d.node0 = node0; d.nb_local = local; d.magic = magic; d.pad = 0;
        descs[stage] = d;
        const uint32_t bytes = (uint32_t)rows * (uint32_t)GF_LD * 4u;
        mbar_arrive_expect_tx(&xfull[stage], bytes);   // release: publishes desc + neighbour words to the aggregators
        bulk_g2s(xring + (size_t)stage * GF_TILE_BYTES, a.x + row0 * GF_LD, bytes, &xfull[stage]);
      }
      if (++stage == GF_XSTAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1 || warp == GF_ISSUER2) {
    // =============================================================================================== MMA issuers
    // A thread issues one tcgen05.mma every ~100 cycles (descriptor arithmetic on the uniform datapath + the election
    // loop around every instruction) while the tensor core needs 48 cycles at N = 64 (scripts/mma_rate2.cu): with one
    // issuer the 48 instructions of a tile took 5 040 cycles and the aggregators sat on `opfree`.  Two warps split the
    // K-blocks (warp 1: 0-1, warp 18: 2-3), each into accumulators of its own that the epilogue adds.
    const int iss = warp == 1 ? 0 : 1;
    if (lane == 0) {
      unsigned i = 0;
      for (long long u = blockIdx.x; u < total; u += gridDim.x, ++i) {
        const uint32_t b = i & 1u, acc = i % GF_ACCS;
        const int rows = __ldg(a.unit_desc + (u % U) * 12 + 3);
        int n16 = (rows + 15) & ~15;
        if (n16 < 16) n16 = 16;
        // D [M = 128 channels, N = n16 tile rows] (+)= A (tensor memory, K-major) * B^T (shared memory, K-major)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        mbar_wait(&opfull[b], (i >> 1) & 1u);
        if (i >= GF_ACCS) mbar_wait(&accfree[acc], ((i / GF_ACCS) - 1u) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem + ((uint32_t)iss * GF_ACCS + acc) * 64u;
        const uint32_t xh = smem_u32(opbuf + (size_t)b * GF_OP_BYTES), xl = xh + 4 * GF_KB_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4 / GF_ISSUERS; ++kk) {
          const int kb = iss * (4 / GF_ISSUERS) + kk;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t wh = tmem + GF_COL_WH + kb * 32 + j * 8, wl = tmem + GF_COL_WL + kb * 32 + j * 8;
            const uint32_t o = kb * GF_KB_BYTES + j * 32;
            gf_mma_ts(tacc, wh, gf_make_desc(xh + o), idesc, (kk | j) ? 1u : 0u);
            gf_mma_ts(tacc, wl, gf_make_desc(xh + o), idesc, 1u);
            gf_mma_ts(tacc, wh, gf_make_desc(xl + o), idesc, 1u);
          }
        }
        gf_commit(&opfree[b]);
        gf_commit(&accdone[acc]);
      }
    }
  } else if (warp < GF_FIRST_EPI) {
    // =============================================================================================== aggregators
    const int cw = warp - GF_FIRST_AGG;
    const int col = lane * 4;                       // this lane's float4 of every row it touches
    const int kb = lane >> 3, c16 = lane & 7;       // its K-block / 16-byte chunk in the operand buffers
    const float one_eps = __fadd_rn(1.0f, a.eps ? __ldg(a.eps) : 0.0f);
    unsigned i = 0;
    for (long long u = blockIdx.x; u < total; u += gridDim.x, ++i) {
      const unsigned xs = i % GF_XSTAGES, b = i & 1u;
      mbar_wait(&xfull[xs], (i / GF_XSTAGES) & 1u);
      if (i >= GF_OPBUFS) mbar_wait(&opfree[b], ((i >> 1) - 1u) & 1u);
      const GfDesc d = descs[xs];
      const float* tile = reinterpret_cast<const float*>(xring + (size_t)xs * GF_TILE_BYTES);
      const uint32_t* nb = nbw + xs * GF_NB_MAXN;
      float* out_tile = a.a_out + d.row0 * (long long)GF_LD;
      uint8_t* oph = opbuf + (size_t)b * GF_OP_BYTES + (size_t)kb * GF_KB_BYTES;
      uint8_t* opl = oph + 4 * GF_KB_BYTES;
#pragma unroll 2
      for (int r = cw; r < d.rows; r += GF_AGG_WARPS) {
        const int slot = (int)__umulhi((unsigned)r, d.magic);
        const int li = r - slot * d.n;
        const float* slot_tile = tile + slot * d.n * GF_LD;
        const float* self_row = tile + r * GF_LD;
        const uint32_t word = d.nb_local ? nb[li] : (GF_NB_SLOW << 24);
        const uint32_t j0 = word & 0xFFu, j1 = (word >> 8) & 0xFFu, j2 = (word >> 16) & 0xFFu, j3 = word >> 24;
        float4 acc;
        if (j3 != GF_NB_SLOW) {
          const float* p0 = j0 == GF_NB_NONE ? zrow : slot_tile + j0 * GF_LD;
          const float* p1 = j1 == GF_NB_NONE ? zrow : slot_tile + j1 * GF_LD;
          const float* p2 = j2 == GF_NB_NONE ? zrow : slot_tile + j2 * GF_LD;
          const float* p3 = j3 == GF_NB_NONE ? zrow : slot_tile + j3 * GF_LD;
          const float4 v0 = gf_lds4(p0 + col), v1 = gf_lds4(p1 + col), v2 = gf_lds4(p2 + col), v3 = gf_lds4(p3 + col);
          acc = gf_add4(gf_add4(gf_add4(gf_add4(make_float4(0.f, 0.f, 0.f, 0.f), v0), v1), v2), v3);
        } else {  // degree > 4 / large graph: walk the CSR in global memory (rows are still read from the tile)
          const int beg = __ldg(a.nbr_ptr + d.node0 + li), end = __ldg(a.nbr_ptr + d.node0 + li + 1);
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int e = beg; e < end; ++e) {
            const int lj = __ldg(a.nbr_idx + e) - d.node0;
            acc = gf_add4(acc, gf_lds4(slot_tile + lj * GF_LD + col));
          }
        }
        const float4 self = gf_lds4(self_row + col);
        acc.x = __fadd_rn(acc.x, __fmul_rn(one_eps, self.x));
        acc.y = __fadd_rn(acc.y, __fmul_rn(one_eps, self.y));
        acc.z = __fadd_rn(acc.z, __fmul_rn(one_eps, self.z));
        acc.w = __fadd_rn(acc.w, __fmul_rn(one_eps, self.w));
        stg4_stream(out_tile + r * GF_LD + col, acc);          // A is kept for the backward's weight gradient
        float4 hh, ll;
        gf_split(acc.x, hh.x, ll.x);
        gf_split(acc.y, hh.y, ll.y);
        gf_split(acc.z, hh.z, ll.z);
        gf_split(acc.w, hh.w, ll.w);
        const uint32_t off = gf_sw128(r, c16);
        *reinterpret_cast<float4*>(oph + off) = hh;
        *reinterpret_cast<float4*>(opl + off) = ll;
      }
      // rows [d.rows, n16) of the operand buffer are never written: the MMA turns them into accumulator COLUMNS the
      // epilogue does not read (columns are independent), so stale values there are harmless.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&xempty[xs]);
        mbar_arrive(&opfull[b]);
      }
    }
  } else {
    // ================================================================================================== epilogue
    // accumulator: TMEM lane = output channel, column = tile row.  Warp e: channels [32 q, 32 q + 32) (q = warp & 3, its
    // TMEM lane quarter), tile rows [32 half, 32 half + 32).
    const int q = warp & 3, half = (warp - GF_FIRST_EPI) >> 2;
    const int n = q * 32 + lane;
    const bool chan = n < a.h, store = n < a.ldh;
    double st_s[2] = {0.0, 0.0}, st_q[2] = {0.0, 0.0};
    unsigned i = 0;
    for (long long u = blockIdx.x; u < total; u += gridDim.x, ++i) {
      const uint32_t acc = i % GF_ACCS;
      const int uu = (int)(u % U), s = (int)(u / U);
      const unsigned rlo = (unsigned)__ldg(a.unit_desc + uu * 12 + 0), rhi = (unsigned)__ldg(a.unit_desc + uu * 12 + 1);
      const int rows = __ldg(a.unit_desc + uu * 12 + 3);
      const long long row0 = (long long)s * a.R + (long long)(((unsigned long long)rhi << 32) | rlo);
      mbar_wait(&accdone[acc], (i / GF_ACCS) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int r0 = half * 32;
      const bool have = r0 < rows;                  // warp-uniform
      uint32_t v[32];
      if (have) {   // the two issuers' partial sums (K-blocks 0-1 and 2-3)
        uint32_t v2[32];
        GF_LD32(v, tmem + acc * 64u + ((uint32_t)(q * 32) << 16) + (uint32_t)r0);
        GF_LD32(v2, tmem + (GF_ACCS + acc) * 64u + ((uint32_t)(q * 32) << 16) + (uint32_t)r0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&accfree[acc]);    // the block is in registers: hand the accumulator back
      if (!have) continue;
      float* hp = a.h_out + (row0 + r0) * (long long)a.ldh + n;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (r0 + j < rows) {                        // warp-uniform
          const float t = __uint_as_float(v[j]);
          if (store) hp[(long long)j * a.ldh] = t;
          s1 += t;
          s2 = fmaf(t, t, s2);
        }
      }
      if (s == 0) { st_s[0] += (double)s1; st_q[0] += (double)s2; }
      else        { st_s[1] += (double)s1; st_q[1] += (double)s2; }
    }
    if (a.stats && chan) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g < a.S) {
          atomicAdd(a.stats + (long long)(g * 2 + 0) * a.h + n, st_s[g]);
          atomicAdd(a.stats + (long long)(g * 2 + 1) * a.h + n, st_q[g]);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static int g_fused = -1;   // -1 undecided (env SB_FUSED_AGG_LINEAR, default on), 0 off, 1 on
extern "C" int sb_set_fused_agg_linear(int32_t enable) {
  const int old = g_fused;
  g_fused = enable ? 1 : 0;
  return old;
}
static bool fused_enabled() {
  if (g_fused < 0) {
    const char* e = getenv("SB_FUSED_AGG_LINEAR");
    g_fused = (e && e[0] == '0') ? 0 : 1;
  }
  return g_fused == 1;
}

// A = aggregate(X), H = A W^T (+ column statistics) in one launch.  Returns SB_ERR_UNSUPPORTED (no error set) for any
// shape outside the fast path; the caller then runs sb_gin_agg followed by sb_linear_fwd (identical results).
extern "C" int sb_gin_linear_fused_fwd(const float* x, float* a_out, float* h_out, double* stats, const float* eps,
                                       const float* w, int64_t w_rs, int64_t w_cs, int32_t K, int32_t h, int64_t ldh,
                                       const int32_t* unit_ptr, const int32_t* unit_desc, const uint32_t* nbr_pack,
                                       const int32_t* nbr_ptr, const int32_t* nbr_idx, int64_t R, int32_t B, int32_t S,
                                       int32_t ld, int32_t tile_rows, int32_t generic, void* stream) {
  if (!fused_enabled() || generic || ld != GF_LD || K != GF_LD || h < 1 || h > 128 || ldh < h || ldh > 128 || S < 1 ||
      S > 2 || tile_rows != GF_ROWS || !unit_desc || !nbr_pack || R <= 0 || B <= 0 || R >= (1ll << 31) ||
      ((uintptr_t)x % 16) || ((uintptr_t)a_out % 16) || x == a_out)
    return SB_ERR_UNSUPPORTED;
  GfArgs a;
  a.x = x; a.a_out = a_out; a.h_out = h_out; a.stats = stats; a.eps = eps; a.w = w; a.w_rs = w_rs; a.w_cs = w_cs;
  a.unit_ptr = unit_ptr; a.unit_desc = unit_desc; a.nbr_pack = nbr_pack; a.nbr_ptr = nbr_ptr; a.nbr_idx = nbr_idx;
  a.R = R; a.B = B; a.S = S; a.h = h; a.ldh = (int)ldh;
  static bool configured = false;
  if (!configured) {
    SB_CUDA(cudaFuncSetAttribute(gin_lin_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GF_DYN_SMEM));
    configured = true;
  }
  gin_lin_fused_kernel<<<sb_num_sms(), GF_THREADS, GF_DYN_SMEM, (cudaStream_t)stream>>>(a);
  SB_CHECK_LAUNCH("sb_gin_linear_fused_fwd");
  return SB_OK;
}

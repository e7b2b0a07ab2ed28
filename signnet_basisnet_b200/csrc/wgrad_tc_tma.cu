// Weight gradient on tcgen05, TMA-fed: the DEFAULT kernel for N == K == 128 and 16-byte aligned rows (the phi stack at
// n_hid = 128); wgrad_tc.cu (register-fed ring) handles every other fast shape.
// Same arithmetic, MMA order, accumulator, epilogue and deterministic two-stage reduction as wgrad_tc_kernel<FAST>; what
// changes is how the operand ring is filled: one thread issues tensor-map TMA loads of the raw 32-row chunks of g and x
// ([32 rows x 32 floats] boxes, SWIZZLE_128B_ATOM_32B = the MN-major tf32 layout the UMMA descriptor names
// SWIZZLE_128B_BASE32B; rows past the end of a group are zero-filled) straight into the head buffers of a 3-stage ring,
// up to three chunks (96 KB) ahead of the MMAs, and the 16 worker warps only add the 3xTF32 tails in place (plus the
// rewritten heads; letting the tensor core truncate the raw fp32 word itself was measured 3 % faster in round 2 but is
// not bit-identical to wgrad_tc.cu and was dropped).
//
// STATUS (round 2, B200): bit-identical to wgrad_tc.cu on every case of scripts/pair_check.cu and tests/test_gpu_wgrad_
// variants.py; 278 us vs 338 us per launch at the phi size of cfg 4 (profiles/r2a_pair_check_mode3.log).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define WM_ROWS 32
#define WM_BLK_BYTES (32 * 128 * 4)   // one [32 rows x 128 features] fp32 operand block = 16 KB
#define WM_STAGE_BYTES (4 * WM_BLK_BYTES)
#define WM_STAGES 3
#define WM_WORKERS 512
#define WM_ISSUERS 2                   // MMA-issuing warps: 16 and 18 (row groups 0-1 / 2-3 of every chunk)
#define WM_THREADS (WM_WORKERS + 96)   // + MMA warp (16) + TMA-issue warp (17) + second MMA warp (18)
#define WM_MAXG 2
#define WM_L2_AHEAD 8   // chunks requested into L2 ahead of the TMA loads (cp.async.bulk.prefetch.L2)

struct WgTmaArgs {
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
};

__device__ __forceinline__ uint64_t wm_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((4096 >> 4) & 0x3FFF) << 16;   // LBO: next 32-feature MN block
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;    // SBO: next atom of 4 rows
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)1 << 61;                        // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void wm_split(float x, float& h, float& l) {
  // == cvt.rna.tf32.f32 for finite x, in two ALU instructions (see tc_split in linear_tc.cu)
  h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  l = x - h;
}
__device__ __forceinline__ void wm_mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void wm_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wm_worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WM_WORKERS) : "memory"); }

__device__ __forceinline__ void wm_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// 3-D tensor-map load (column, row inside the group, group) of one [32 rows x 32 floats] box
__device__ __forceinline__ void wm_tma_load(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(WM_THREADS, 1)
wgrad_tc_tma_kernel(const WgTmaArgs a, const __grid_constant__ CUtensorMap tmg, const __grid_constant__ CUtensorMap tmx) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;   // [WM_STAGES][g_hi | g_lo | x_hi | x_lo]
  __shared__ float s_pa[WM_MAXG * 128], s_pc[WM_MAXG * 128];
  __shared__ uint64_t tma_full[WM_STAGES], full[WM_STAGES], mma_done[WM_STAGES], acc_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K;

  for (int idx = tid; idx < WM_MAXG * 128; idx += WM_THREADS) {
    const int g = idx >> 7, c = idx & 127;
    const bool ok = a.pro && g < a.G && c < K;
    s_pa[idx] = ok ? __ldg(a.pa + (long long)g * K + c) : 1.f;
    s_pc[idx] = ok ? __ldg(a.pc + (long long)g * K + c) : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < WM_STAGES; ++s) {
      mbar_init(&tma_full[s], 1);
      mbar_init(&full[s], WM_WORKERS / 32);
      mbar_init(&mma_done[s], WM_ISSUERS);   // one tcgen05.commit per issuing warp
    }
    mbar_init(&acc_done, WM_ISSUERS);
    mbar_fence_init();
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128 * WM_ISSUERS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const long long cpg = (a.R + WM_ROWS - 1) / WM_ROWS;
  const long long nch = cpg * a.G;

  if (warp == 16 || warp == 18) {
    // ============================================================================================== MMA issuers
    // one thread issues a tcgen05.mma every ~100 cycles (descriptor arithmetic + election loop) while the tensor core
    // needs 64 (scripts/mma_rate2.cu): 12 instructions per 32-row chunk from one thread = 1 260 cycles next to the 1 400
    // cycles the chunk's 32 KB take to arrive from HBM.  Two warps split the chunk's four 8-row groups and accumulate into
    // tensor-memory buffers of their own; the epilogue (once per CTA) adds them.
    const int iss = warp == 16 ? 0 : 1;
    if (lane == 0) {
      // D[M = 128 (n), N = KP (k)] += A^T-major G chunk x X chunk, both MN-major tf32
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(a.KP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t tacc = tmem + (uint32_t)iss * 128u;
      unsigned cnt = 0;
      for (long long c = blockIdx.x; c < nch; c += gridDim.x, ++cnt) {
        const int stage = cnt % WM_STAGES;
        mbar_wait(&full[stage], (cnt / WM_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t gh = smem_u32(ring + stage * WM_STAGE_BYTES), gl = gh + WM_BLK_BYTES;
        const uint32_t xh = gl + WM_BLK_BYTES, xl = xh + WM_BLK_BYTES;
#pragma unroll
        for (int jj = 0; jj < 4 / WM_ISSUERS; ++jj) {   // this issuer's groups of 8 rows
          const uint32_t o = (uint32_t)(iss * (4 / WM_ISSUERS) + jj) * 1024;
          wm_mma(tacc, wm_make_desc(gh + o), wm_make_desc(xh + o), idesc, (cnt | jj) ? 1u : 0u);
          wm_mma(tacc, wm_make_desc(gh + o), wm_make_desc(xl + o), idesc, 1u);
          wm_mma(tacc, wm_make_desc(gl + o), wm_make_desc(xh + o), idesc, 1u);
        }
        wm_commit(&mma_done[stage]);
      }
      wm_commit(&acc_done);
    }
  } else if (warp == 17) {
    // =============================================================================================== TMA issuer
    if (lane == 0) {
      auto l2_chunk = [&](long long cc) {   // one bulk prefetch per operand: a chunk's rows are contiguous
        const int g = (cc >= cpg) ? 1 : 0;
        const long long row0 = (cc - (long long)g * cpg) * WM_ROWS;
        const int rows = (int)((a.R - row0 < WM_ROWS) ? (a.R - row0) : WM_ROWS);
        const long long base = (long long)g * a.R + row0;
        wm_prefetch_l2(a.g + base * a.ldg, (uint32_t)(rows * a.ldg * 4));
        wm_prefetch_l2(a.x + base * a.ldx, (uint32_t)(rows * a.ldx * 4));
      };
      for (int s = WM_STAGES; s < WM_L2_AHEAD; ++s)
        if (blockIdx.x + (long long)s * gridDim.x < nch) l2_chunk(blockIdx.x + (long long)s * gridDim.x);
      unsigned cnt = 0;
      for (long long c = blockIdx.x; c < nch; c += gridDim.x, ++cnt) {
        const int stage = cnt % WM_STAGES;
        const unsigned use = cnt / WM_STAGES;
        if (use > 0) mbar_wait(&mma_done[stage], (use - 1) & 1);
        const int g = (c >= cpg) ? 1 : 0;
        const int row0 = (int)((c - (long long)g * cpg) * WM_ROWS);
        uint8_t* sb = ring + stage * WM_STAGE_BYTES;
        mbar_arrive_expect_tx(&tma_full[stage], 2 * WM_BLK_BYTES);
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {   // heads of g and x: [blk][32 rows][128 B]
          wm_tma_load(sb + blk * 4096, &tmg, blk * 32, row0, g, &tma_full[stage]);
          wm_tma_load(sb + 2 * WM_BLK_BYTES + blk * 4096, &tmx, blk * 32, row0, g, &tma_full[stage]);
        }
        const long long cl = c + (long long)WM_L2_AHEAD * gridDim.x;
        if (cl < nch) l2_chunk(cl);
      }
    }
  } else {
    // ================================================================================================== workers
    // item q of a thread: idx = tid + 512 q -> 32-feature block (idx >> 8), row (idx >> 3) & 31, float4 (idx & 7)
    float dbs[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const int row = (tid >> 3) & 31, c4 = tid & 7;
    unsigned cnt = 0;
    for (long long c = blockIdx.x; c < nch; c += gridDim.x, ++cnt) {
      const int stage = cnt % WM_STAGES;
      mbar_wait(&tma_full[stage], (cnt / WM_STAGES) & 1);
      // (the tail buffers are free: the TMA warp saw mma_done of this stage's previous use before it started the copies)
      const int g = (c >= cpg) ? 1 : 0;
      const long long row0 = (c - (long long)g * cpg) * WM_ROWS;
      const int rows = (int)((a.R - row0 < WM_ROWS) ? (a.R - row0) : WM_ROWS);
      uint8_t* sb = ring + stage * WM_STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int blk = (tid + WM_WORKERS * q) >> 8;
        const int col = blk * 32 + c4 * 4;
        const uint32_t off = (uint32_t)(blk * 4096 + row * 128 + (((c4 >> 1) ^ (row & 3)) << 5) + (c4 & 1) * 16);
        // ---- g: never has a prologue; rows past the end of the group were zero-filled by the TMA unit
        const float4 vg = *reinterpret_cast<const float4*>(sb + off);
        dbs[q][0] += vg.x; dbs[q][1] += vg.y; dbs[q][2] += vg.z; dbs[q][3] += vg.w;
        float4 h, l;
        wm_split(vg.x, h.x, l.x); wm_split(vg.y, h.y, l.y); wm_split(vg.z, h.z, l.z); wm_split(vg.w, h.w, l.w);
        *reinterpret_cast<float4*>(sb + off) = h;
        *reinterpret_cast<float4*>(sb + WM_BLK_BYTES + off) = l;
        // ---- x: the forward prologue (BatchNorm affine / ReLU) is recomputed, never stored
        const float4 vx = *reinterpret_cast<const float4*>(sb + 2 * WM_BLK_BYTES + off);
        {
          float tx[4] = {vx.x, vx.y, vx.z, vx.w};
          if (a.pro) {
            const float4 pa4 = *reinterpret_cast<const float4*>(&s_pa[g * 128 + col]);
            const float4 pc4 = *reinterpret_cast<const float4*>(&s_pc[g * 128 + col]);
            tx[0] = fmaf(pa4.x, tx[0], pc4.x); tx[1] = fmaf(pa4.y, tx[1], pc4.y);
            tx[2] = fmaf(pa4.z, tx[2], pc4.z); tx[3] = fmaf(pa4.w, tx[3], pc4.w);
            if (a.pro == 2) {
              tx[0] = fmaxf(tx[0], 0.f); tx[1] = fmaxf(tx[1], 0.f); tx[2] = fmaxf(tx[2], 0.f); tx[3] = fmaxf(tx[3], 0.f);
            }
            if (!(row < rows)) tx[0] = tx[1] = tx[2] = tx[3] = 0.f;   // zero-filled rows must not pick up the shift
          }
          wm_split(tx[0], h.x, l.x); wm_split(tx[1], h.y, l.y); wm_split(tx[2], h.z, l.z); wm_split(tx[3], h.w, l.w);
          *reinterpret_cast<float4*>(sb + 2 * WM_BLK_BYTES + off) = h;
        }
        *reinterpret_cast<float4*>(sb + 3 * WM_BLK_BYTES + off) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);   // one arrival per worker warp: no block-wide barrier per chunk
    }

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
        uint32_t v2[32];   // the second issuer's accumulator
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
                     "%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v2[0]), "=r"(v2[1]), "=r"(v2[2]), "=r"(v2[3]), "=r"(v2[4]), "=r"(v2[5]), "=r"(v2[6]), "=r"(v2[7]),
                       "=r"(v2[8]), "=r"(v2[9]), "=r"(v2[10]), "=r"(v2[11]), "=r"(v2[12]), "=r"(v2[13]), "=r"(v2[14]), "=r"(v2[15]),
                       "=r"(v2[16]), "=r"(v2[17]), "=r"(v2[18]), "=r"(v2[19]), "=r"(v2[20]), "=r"(v2[21]), "=r"(v2[22]), "=r"(v2[23]),
                       "=r"(v2[24]), "=r"(v2[25]), "=r"(v2[26]), "=r"(v2[27]), "=r"(v2[28]), "=r"(v2[29]), "=r"(v2[30]), "=r"(v2[31])
                     : "r"(tmem + 128u + ((uint32_t)(q * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
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
      wm_worker_sync();
      float* red = reinterpret_cast<float*>(ring);   // [32 row slots][128 columns]
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int idx = tid + WM_WORKERS * q;
        const int rslot = (idx >> 3) & 31, col = (idx >> 8) * 32 + (idx & 7) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) red[rslot * 128 + col + j] = dbs[q][j];
      }
      wm_worker_sync();
      if (tid < 128) {
        float s = 0.f;
        for (int r = 0; r < 32; ++r) s += red[r * 128 + tid];
        a.part_b[(size_t)blockIdx.x * 128 + tid] = s;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128 * WM_ISSUERS));
}

int sb_wgrad_reduce_launch(const float* part_w, const float* part_b, int nparts, int BN, int BK, int N, int K, float* dw,
                           long long rs, long long cs, float* db, int accumulate, cudaStream_t st);

// ---- host side: tensor maps through the driver entry point (no link-time dependency on libcuda) ------------------
typedef CUresult (*wm_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// tuning knob of the opt-in kernels: SB_TMA_L2PROMO = 0 (none, default) | 1 (64 B) | 2 (128 B) | 3 (256 B)
static CUtensorMapL2promotion wm_l2promo() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SB_TMA_L2PROMO");
    v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
         : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
}
static wm_encode_fn wm_encoder() {
  static wm_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (wm_encode_fn)p;
  }
  return fn;
}
// [G][R][128] fp32 view with row stride ld (floats); box {32 floats, 32 rows, 1}; 32-byte-atom 128-byte swizzle
static int wm_make_map(CUtensorMap* tm, const float* base, int64_t ld, int64_t R, int32_t G) {
  wm_encode_fn enc = wm_encoder();
  if (!enc) return SB_ERR_UNSUPPORTED;
  const cuuint64_t dims[3] = {128, (cuuint64_t)R, (cuuint64_t)G};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)R * (cuuint64_t)ld * 4};
  const cuuint32_t box[3] = {32, WM_ROWS, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         wm_l2promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (r == CUDA_SUCCESS) ? SB_OK : SB_ERR_UNSUPPORTED;
}

// Returns SB_ERR_UNSUPPORTED (without setting an error) for anything but N == K == 128 with 16-byte aligned rows; the
// caller then uses wgrad_tc_kernel (same contract).
int sb_wgrad_tc_tma_launch(const float* gy, int64_t ldg, const float* x, int64_t ldx, int64_t R, int32_t G, int32_t N,
                           int32_t K, int32_t pro, const float* pa, const float* pc, float* dw, int64_t dw_rs,
                           int64_t dw_cs, float* db, int32_t accumulate, float* workspace, cudaStream_t st) {
  const bool gvec = (ldg % 4 == 0) && ((uintptr_t)gy % 16 == 0);
  const bool xvec = (ldx % 4 == 0) && ((uintptr_t)x % 16 == 0);
  if (K != 128 || N != 128 || G > WM_MAXG || R * G < 4096 || !gvec || !xvec || R >= (1ll << 31)) return SB_ERR_UNSUPPORTED;
  CUtensorMap tmg, tmx;
  if (wm_make_map(&tmg, gy, ldg, R, G) != SB_OK || wm_make_map(&tmx, x, ldx, R, G) != SB_OK) return SB_ERR_UNSUPPORTED;
  WgTmaArgs a;
  a.g = gy; a.ldg = ldg; a.x = x; a.ldx = ldx; a.R = R; a.G = G; a.N = N; a.K = K; a.KP = 128;
  a.pro = pro; a.pa = pa; a.pc = pc;
  const long long nch = sb_ceil_div(R, WM_ROWS) * G;
  long long grid = sb_num_sms();
  if (grid > nch) grid = nch;
  a.part_w = workspace;
  a.part_b = db ? workspace + (size_t)grid * 128 * 128 : nullptr;
  const size_t smem = (size_t)WM_STAGES * WM_STAGE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    SB_CUDA(cudaFuncSetAttribute(wgrad_tc_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  wgrad_tc_tma_kernel<<<(unsigned)grid, WM_THREADS, smem, st>>>(a, tmg, tmx);
  SB_CHECK_LAUNCH("sb_linear_wgrad(tcgen05 + TMA)");
  return sb_wgrad_reduce_launch(a.part_w, a.part_b, (int)grid, 128, 128, N, K, dw, dw_rs, dw_cs, db, accumulate, st);
}

// K2 on the 5th-generation tensor cores: y = f(x) W^T (+ bias, relu, column statistics) with tcgen05.mma kind::tf32
// and error-compensated operands (3xTF32):  x = xh + xl, w = wh + wl with xh/wh the tf32-rounded heads, and
//     x w  ~=  xh wh + xh wl + xl wh        (fp32 accumulation in tensor memory)
// which keeps the result within ~1e-6 relative of an fp32 FFMA contraction — BASELINE.json's 1e-5 parity bar rules out
// a single-pass TF32/bf16 product.  Same contract as linear_fwd_kernel (linear.cu): prologue none | BN-affine |
// BN-affine+ReLU applied once per element, epilogue bias / ReLU / per-(group, channel) fp64 column sums.
//
// One persistent CTA per SM, three warp roles running concurrently (no block-wide barrier inside the tile loop):
//   warps 0..7   producers: stream the input as [128 rows x 32 floats] K-blocks (register prefetch two K-blocks ahead,
//                bulk L2 prefetch three tiles ahead), apply the prologue, split head/tail and st.shared them swizzled
//                into a 2-stage operand ring; each warp arrives on full[stage] on its own.
//   warp 16      MMA issuer: waits full[stage], issues 12 x UMMA 128 x N x 8 (kind::tf32) per K-block against the
//                resident weight (head and tail, canonical K-major SWIZZLE_128B, 128 KB); tcgen05.commit recycles the
//                ring stage (mma_done) and publishes the accumulator (acc_done).  Two accumulators in tensor memory
//                alternate, so the MMAs of tile i+1 run while tile i is drained (acc_free hands a buffer back).
//   warps 8..15  epilogue: tcgen05.ld of a [32 lanes x 32 columns] block (two per warp and tile) -> bias/ReLU ->
//                per-warp XOR-swizzled shared tile -> 128-bit stores of whole 128-byte row segments + per-column
//                BatchNorm partial sums (fp32 over the block's 32 rows, fp64 across tiles, one atomic per kernel).
// The first version ran split -> MMA -> staged epilogue one after the other on the same 16 warps behind bar.syncs and
// sat at 24 % of DRAM bandwidth with the tensor pipe 19 % busy (profiles/r1g_linear_tc_ncu.csv).
#include "common.cuh"
#include "../../include/signnet_b200.h"

#define TC_BM 128
#define TC_KB 32                      // floats per K-block = one 128-byte swizzle row
#define TC_BLK_BYTES (128 * 128)      // one [128 rows x 32 floats] operand block
#define TC_WORKERS 512
#define TC_PROD 256                    // producer threads (warps 0..7); warps 8..15 are the epilogue
#define TC_THREADS (TC_WORKERS + 32)
#define TC_MAXG 2
static_assert(TC_MAXG == 2, "the tile -> group mapping below assumes at most two groups");
#define TC_ESTAGE_BYTES (8 * 32 * 32 * 4)    // per-epilogue-warp [32 rows][32 cols] fp32 transposition tile
#define TC_L2_AHEAD 2   // tiles requested into L2 ahead of the one-tile register prefetch

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
  // == cvt.rna.tf32.f32 for finite x (adding half an ulp of the 10-bit mantissa to the magnitude bits, then truncating),
  // in two ALU instructions instead of the five the conversion expands to
  h = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
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

// Column sums over the 32 lanes of a warp: lane l ends up with sum_lanes v[l] (a transposing butterfly, 31 shuffles).
__device__ __forceinline__ float tc_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

// FAST: K and N multiples of 32, 16-byte aligned rows on both sides, ycols == N (accumulate only without ReLU / statistics) -> every per-element
// bounds / alignment predicate of the generic path disappears at compile time (it was ~60 % of the issued instructions).
// ACC (FAST only): y += x W^T - a separate instantiation, so the plain kernel's register allocation is untouched (sharing
// one kernel spilled 176 bytes and took the phi-size launch from 282 to 366 us).
template <bool FAST, bool ACC = false>
__global__ void __launch_bounds__(TC_THREADS, 1) linear_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = a.nkb;
  uint8_t* Wh = smem;                              // [nkb][16 KB]
  uint8_t* Wl = Wh + nkb * TC_BLK_BYTES;
  uint8_t* ring = Wl + nkb * TC_BLK_BYTES;         // 2 stages x (head 16 KB | tail 16 KB) = 64 KB
  float* estage = reinterpret_cast<float*>(ring + 4 * TC_BLK_BYTES);   // [8 warps][32 rows][32 cols] epilogue staging
  __shared__ __align__(16) float s_pa[TC_MAXG * 128], s_pc[TC_MAXG * 128], s_bias[128];
  __shared__ uint64_t full[2], mma_done[2], acc_done[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, N = a.N;

  // ---- one-time: split the weight into head/tail, canonical K-major SW128 blocks; rows >= N and columns >= K are 0
  // lanes run along k for either weight orientation: one 128-byte swizzle row per warp store = conflict-free.  (With
  // lanes along n - the coalesced order for the transposed weight of the input-gradient launches - all 32 stores of a
  // warp hit the same bank: 32-way conflicts cost ~10 % of such a launch; the strided global reads are L2 hits.)
  // Six loads per thread are issued before their first use: the prologue is ~10 us of a 46 us launch at 139 k rows (one
  // eighth of the phi size) when every load waits for the previous one's store.
  {
    constexpr int UN = 6;
    const int total = nkb * 128 * TC_KB, rowlen = nkb * TC_KB;
    for (int base0 = tid; base0 < total; base0 += TC_THREADS * UN) {
      float v[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int idx = base0 + u * TC_THREADS;
        const int n = idx / rowlen, k = idx - n * rowlen;
        v[u] = (idx < total && k < K && n < N) ? __ldg(a.w + (long long)n * a.w_rs + (long long)k * a.w_cs) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int idx = base0 + u * TC_THREADS;
        if (idx < total) {
          const int n = idx / rowlen, k = idx - n * rowlen;
          float h, l;
          tc_split(v[u], h, l);
          const uint32_t off = (uint32_t)(k / TC_KB) * TC_BLK_BYTES + tc_sw128(n, (k % TC_KB) >> 2) + (uint32_t)(k & 3) * 4;
          *reinterpret_cast<float*>(Wh + off) = h;
          *reinterpret_cast<float*>(Wl + off) = l;
        }
      }
    }
  }
  for (int idx = tid; idx < TC_MAXG * 128; idx += TC_THREADS) {
    const int g = idx >> 7, c = idx & 127;
    const bool ok = a.pro && g < a.G && c < K;
    s_pa[idx] = ok ? __ldg(a.pa + (long long)g * K + c) : 1.f;
    s_pc[idx] = ok ? __ldg(a.pc + (long long)g * K + c) : 0.f;
  }
  for (int idx = tid; idx < 128; idx += TC_THREADS) s_bias[idx] = (a.bias && idx < N) ? __ldg(a.bias + idx) : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], TC_PROD / 32);          // one arrival per producer warp
      mbar_init(&mma_done[i], 1);
      mbar_init(&acc_done[i], 1);
      mbar_init(&acc_free[i], (TC_WORKERS - TC_PROD) / 32);   // one arrival per epilogue warp
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

  const long long tpg = (a.R + TC_BM - 1) / TC_BM;
  const long long ntiles = tpg * a.G;

  if (warp == 16) {
    // =============================================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.NP >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      unsigned cnt = 0, ti = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        const uint32_t buf = ti & 1u;
        if (ti >= 2) {   // the workers have drained the accumulator this tile reuses
          mbar_wait(&acc_free[buf], ((ti >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tacc = tmem + buf * 128u;
        for (int kb = 0; kb < nkb; ++kb, ++cnt) {
          const int stage = cnt & 1;
          mbar_wait(&full[stage], (cnt >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = smem_u32(ring + stage * 2 * TC_BLK_BYTES), al = ah + TC_BLK_BYTES;
          const uint32_t wh = smem_u32(Wh + kb * TC_BLK_BYTES), wl = smem_u32(Wl + kb * TC_BLK_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t o = j * 32;
            tc_mma(tacc, tc_make_desc(ah + o), tc_make_desc(wh + o), idesc, (kb | j) ? 1u : 0u);
            tc_mma(tacc, tc_make_desc(ah + o), tc_make_desc(wl + o), idesc, 1u);
            tc_mma(tacc, tc_make_desc(al + o), tc_make_desc(wh + o), idesc, 1u);
          }
          tc_commit(&mma_done[stage]);
          if (kb == nkb - 1) tc_commit(&acc_done[buf]);
        }
      }
    }
  } else if (warp < TC_PROD / 32) {
    // ================================================================================================ producers
    // thread -> float4 column c4 of rows (tid >> 3) + 32 q, q < 4, of every K-block
    const int prow = tid >> 3, c4 = tid & 7;
    const float* xthread = a.x + (long long)prow * a.ldx + c4 * 4;   // this thread's element of row 0, K-block 0
    const long long xstride32 = 32 * a.ldx;
    float4 pre[2][4];                                  // two K-blocks in flight (slot = position parity)
    struct Cur { long long tile; int kb; };
    auto advance = [&](Cur& c) {
      if (++c.kb == nkb) { c.kb = 0; c.tile += gridDim.x; }
    };
#define TC_LOAD_BLOCK(CUR, SLOT)                                                                    \
    if ((CUR).tile < ntiles) {                                                                      \
      const int g_ = ((CUR).tile >= tpg) ? 1 : 0;                                                      \
      const long long row0_ = ((CUR).tile - (long long)g_ * tpg) * TC_BM;                           \
      const int rows_ = (int)((a.R - row0_ < TC_BM) ? (a.R - row0_) : TC_BM);                       \
      const long long base_ = (long long)g_ * a.R + row0_;                                          \
      const int col_ = (CUR).kb * TC_KB + c4 * 4;                                                   \
      const float* pb_ = xthread + base_ * a.ldx + (CUR).kb * TC_KB;                                \
      _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                               \
        const int row_ = prow + 32 * q;                                                             \
        float4 v_ = make_float4(0.f, 0.f, 0.f, 0.f);                                                \
        if (FAST) {                                                                                 \
          if (row_ < rows_) v_ = ldg4(pb_ + q * xstride32);                                         \
        } else if (row_ < rows_ && col_ < K) {                                                      \
          const float* p_ = pb_ + q * xstride32;                                                    \
          if (a.xvec) {                                                                             \
            v_ = ldg4(p_);                                                                          \
          } else {                                                                                  \
            v_.x = __ldg(p_);                                                                       \
            if (col_ + 1 < K) v_.y = __ldg(p_ + 1);                                                 \
            if (col_ + 2 < K) v_.z = __ldg(p_ + 2);                                                 \
            if (col_ + 3 < K) v_.w = __ldg(p_ + 3);                                                 \
          }                                                                                         \
        }                                                                                           \
        pre[SLOT][q] = v_;                                                                          \
      }                                                                                             \
    }
    // prologue (BatchNorm affine / ReLU) + zeroing of dead rows and the K tail + head/tail split of one K-block
    auto store_block = [&](int stage, const Cur& c, const float4 (&pv)[4]) {
      const int g = (c.tile >= tpg) ? 1 : 0;   // G <= TC_MAXG = 2: no 64-bit division in the hot loop
      const long long row0 = (c.tile - (long long)g * tpg) * TC_BM;
      const int rows = (int)((a.R - row0 < TC_BM) ? (a.R - row0) : TC_BM);
      uint8_t* sh = ring + stage * 2 * TC_BLK_BYTES;
      uint8_t* sl = sh + TC_BLK_BYTES;
      const int col = c.kb * TC_KB + c4 * 4;
      float4 pa4 = make_float4(1.f, 1.f, 1.f, 1.f), pc4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.pro) {   // coefficients of columns >= K are (1, 0)
        pa4 = *reinterpret_cast<const float4*>(&s_pa[g * 128 + col]);
        pc4 = *reinterpret_cast<const float4*>(&s_pc[g * 128 + col]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int row = prow + 32 * q;
        float t[4] = {pv[q].x, pv[q].y, pv[q].z, pv[q].w};
        const bool live = row < rows;
        if (a.pro) {
          t[0] = fmaf(pa4.x, t[0], pc4.x); t[1] = fmaf(pa4.y, t[1], pc4.y);
          t[2] = fmaf(pa4.z, t[2], pc4.z); t[3] = fmaf(pa4.w, t[3], pc4.w);
          if (a.pro == 2) {
            t[0] = fmaxf(t[0], 0.f); t[1] = fmaxf(t[1], 0.f); t[2] = fmaxf(t[2], 0.f); t[3] = fmaxf(t[3], 0.f);
          }
        }
        if (FAST) {
          if (!live) t[0] = t[1] = t[2] = t[3] = 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (!(live && col + j < K)) t[j] = 0.f;
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
    auto l2_ahead = [&](long long tile) {   // one bulk prefetch per tile: its rows are contiguous
      const long long pt = tile + (long long)TC_L2_AHEAD * gridDim.x;
      if (pt < ntiles) {
        const int pg = (pt >= tpg) ? 1 : 0;
        const long long prow0 = (pt - (long long)pg * tpg) * TC_BM;
        const int prows = (int)((a.R - prow0 < TC_BM) ? (a.R - prow0) : TC_BM);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.x + ((long long)pg * a.R + prow0) * a.ldx),
                     "r"((uint32_t)(prows * a.ldx * 4)) : "memory");
      }
    };
    auto publish = [&](int stage) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);
    };

    Cur cur{(long long)blockIdx.x, 0}, nxt = cur;
    TC_LOAD_BLOCK(nxt, 0)
    advance(nxt);
    TC_LOAD_BLOCK(nxt, 1)
    advance(nxt);
    unsigned cnt = 0;
    while (cur.tile < ntiles) {
      // position cnt (even) -> stage 0 / slot 0
      if (cnt >= 2) mbar_wait(&mma_done[0], ((cnt >> 1) - 1) & 1);   // MMAs that read this stage have retired
      if (tid == 0 && cur.kb == 0 && (FAST || a.xvec)) l2_ahead(cur.tile);
      store_block(0, cur, pre[0]);
      publish(0);
      TC_LOAD_BLOCK(nxt, 0)
      advance(nxt);
      advance(cur);
      ++cnt;
      if (cur.tile >= ntiles) break;
      // position cnt (odd) -> stage 1 / slot 1
      if (cnt >= 2) mbar_wait(&mma_done[1], ((cnt >> 1) - 1) & 1);
      if (tid == 0 && cur.kb == 0 && (FAST || a.xvec)) l2_ahead(cur.tile);
      store_block(1, cur, pre[1]);
      publish(1);
      TC_LOAD_BLOCK(nxt, 1)
      advance(nxt);
      advance(cur);
      ++cnt;
    }
#undef TC_LOAD_BLOCK
  } else {
    // ================================================================================================= epilogue
    // warp e owns TMEM lanes [32 q, 32 q + 32) and the column blocks cb = (e >> 2) and (e >> 2) + 2
    const int e = warp - TC_PROD / 32, eq = e & 3;
    float* wst = estage + e * (32 * 32);               // [32 rows][32 cols], 16-byte chunks XOR-swizzled by (row & 7)
    const long long ystride4 = 4 * a.ldy;
    double st_s[2][TC_MAXG] = {{0.0, 0.0}, {0.0, 0.0}}, st_q[2][TC_MAXG] = {{0.0, 0.0}, {0.0, 0.0}};
    unsigned ti = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int g = (tile >= tpg) ? 1 : 0;
      const long long row0 = (tile - (long long)g * tpg) * TC_BM;
      const int rows = (int)((a.R - row0 < TC_BM) ? (a.R - row0) : TC_BM);
      const long long base = (long long)g * a.R + row0;
      const uint32_t buf = ti & 1u;
      mbar_wait(&acc_done[buf], (ti >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = eq * 32 + lane;
      const bool live = row < rows;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int ec0 = ((e >> 2) + 2 * hh) * 32;
        uint32_t v[32];
        if (ec0 < a.NP) {
          TC_LD32(v, tmem + buf * 128u + ((uint32_t)(eq * 32) << 16) + (uint32_t)ec0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (hh == 1) {   // both blocks are in registers / done: hand the accumulator back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_free[buf]);
        }
        if (ec0 >= a.NP) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[ec0 + i * 4]);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = ec0 + i * 4 + j;
            float t = __uint_as_float(v[i * 4 + j]) + bb[j];
            if (!FAST && a.accumulate && live && col < N) t += a.y[(base + row) * a.ldy + col];
            if (a.relu) t = fmaxf(t, 0.f);
            o[j] = ((FAST || col < N) && live) ? t : 0.f;
          }
          *reinterpret_cast<float4*>(wst + lane * 32 + ((i ^ (lane & 7)) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        // coalesced write-back: 4 rows x 128 contiguous bytes per instruction
        {
          const int c = lane & 7;
          float* ybase = a.y + (base + eq * 32 + (lane >> 3)) * a.ldy + ec0 + c * 4;
          float4 yold[8];   // FAST accumulate: all eight loads of the old y in flight before the first dependent store
          if (ACC) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              yold[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (eq * 32 + it * 4 + (lane >> 3) < rows) yold[it] = *reinterpret_cast<const float4*>(ybase + it * ystride4);
            }
          }
          if (FAST || ec0 + c * 4 < a.ycols) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + (lane >> 3);
              const int grow = eq * 32 + r;
              if (grow < rows) {
                float4 o4 = *reinterpret_cast<const float4*>(wst + r * 32 + ((c ^ (r & 7)) << 2));
                float* yp = ybase + it * ystride4;
                if (ACC) {   // y += x W^T: the old values arrived as the same coalesced 128-byte segments
                  o4.x += yold[it].x; o4.y += yold[it].y; o4.z += yold[it].z; o4.w += yold[it].w;
                }
                if (FAST || (a.yvec && ec0 + c * 4 + 3 < a.ycols)) {
                  *reinterpret_cast<float4*>(yp) = o4;
                } else {
                  const float t4[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (ec0 + c * 4 + j < a.ycols) yp[j] = t4[j];
                }
              }
            }
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
    if (a.stats) {   // 4 warps (lane quarters) per column, once per kernel
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = ((e >> 2) + 2 * hh) * 32 + lane;
        if (col < N) {
#pragma unroll
          for (int g = 0; g < TC_MAXG; ++g) {
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
  const size_t smem = (size_t)2 * a.nkb * TC_BLK_BYTES + 4 * TC_BLK_BYTES + TC_ESTAGE_BYTES;
  static bool configured = false;
  if (!configured) {
    const int mx = 2 * 4 * TC_BLK_BYTES + 4 * TC_BLK_BYTES + TC_ESTAGE_BYTES;
    SB_CUDA(cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    SB_CUDA(cudaFuncSetAttribute((linear_tc_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    SB_CUDA(cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    configured = true;
  }
  const long long ntiles = sb_ceil_div(R, TC_BM) * G;
  long long grid = sb_num_sms();
  if (grid > ntiles) grid = ntiles;
  // FAST with accumulate: the epilogue adds the old y in its write-back stage - after ReLU and the column statistics
  // would have been taken, so those two combinations stay on the generic template
  const bool fast = a.xvec && a.yvec && (K % 32 == 0) && (N % 32 == 0) && a.ycols == N && !(accumulate && (relu || stats));
  if (fast && accumulate) linear_tc_kernel<true, true><<<(unsigned)grid, TC_THREADS, smem, st>>>(a);
  else if (fast) linear_tc_kernel<true><<<(unsigned)grid, TC_THREADS, smem, st>>>(a);
  else linear_tc_kernel<false><<<(unsigned)grid, TC_THREADS, smem, st>>>(a);
  SB_CHECK_LAUNCH("sb_linear_fwd(tcgen05)");
  return SB_OK;
}

// K7 tensor-core path — self-attention of a node over its k_b <= 40 eigenvector-slot tokens, d_k = 32, one WARP per
// (node, head), every contraction (Q K^T, P V and the five of the backward) on the warp-level tensor-core instruction
// mma.sync.m16n8k8 tf32 with error compensation (x = hi + lo; lo*hi + hi*lo + hi*hi, fp32 accumulate: the 1e-5 bar of
// the path does not survive single-pass TF32).  Semantics identical to the FFMA kernels in attention_fast.cu and the
// generic ones in transformer.cu (Alchemy/sign_net/model_utils/transformer_module.py:44-58,76-102).
//
// Why the warp-level instruction and not tcgen05: a token set has 9-37 rows; a 128-row tcgen05 tile would hold 3-5 nodes
// with a block-diagonal mask (13 % useful).  m16n8k8 tiles waste at most 15 rows and need neither tensor memory nor
// descriptors.
//
// No shared-memory staging: the contraction index of an MMA can be permuted freely as long as both operands agree, and
// so can the output-column index, so every fragment is loaded from (or stored to) global memory as 128-bit pieces of a
// 128-byte head row:
//   row pattern    (A of Q K^T, and its B = K):  lane (g, t) holds columns {4t..4t+3, 16+4t..16+4t+3} of row g (+8);
//                  k-step kk contracts the lane's values 2kk and 2kk+1;
//   column pattern (B = V of P V):               lane (g, t) holds V[8kt + 2t (+1)][coff(g)..coff(g)+3],
//                  coff(g) = 16 (g & 1) + 4 (g >> 1); the accumulator of output tile ntd then holds columns
//                  4t + ntd and 16 + 4t + ntd, i.e. the lane ends up with the same 8 columns of its rows as it loads;
//   the score accumulators (row g, keys 8nt + 2t, +1) are used directly as the A fragment of P V with the key index
//   permuted (k = t <-> key 8nt + 2t, k = t+4 <-> key 8nt + 2t + 1): no shuffles, no transposition.
// The backward needs P^T and dS^T for dV and dK: phase A (rows = queries) writes Pdrop and dS to shared memory
// [query][key] (2 x 48 x 44 floats per warp), phase B (rows = keys) reads them back transposed as its A fragments.
#include "attention.cuh"
#include "../../include/signnet_b200.h"

#define AM_KMAX 40
#define AM_NT 5        // key tiles of 8
#define AM_WARPS 4

namespace {

// round to tf32 (10 mantissa bits), ties away from zero - what cvt.rna.tf32.f32 computes for finite values, in two integer
// instructions (the cvt expands to ~6 with its NaN / infinity handling; the values here are finite activations)
__device__ __forceinline__ uint32_t tf32_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// x = hi + lo: hi = x rounded to tf32 (10 mantissa bits), lo = x - hi exactly (|lo| <= 2^-12 |x|).  lo goes to the tensor
// core as it is: the instruction ignores the 13 low mantissa bits, an error <= 2^-22 |x| - the size of the lo*lo product
// the scheme drops anyway - for two instructions less per value than rounding it.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rna(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
struct Frag8 {   // 8 values of one row (row pattern), split
  uint32_t hi[8], lo[8];
};
// columns {4t..4t+3, 16+4t..16+4t+3} of a 32-float head row; zeros for a row outside the token set.  Loading (raw) and
// splitting are separate steps so that every load of a phase is in flight before the first dependent instruction.
struct Raw8 {
  float4 x, y;
};
__device__ __forceinline__ Raw8 load_raw(const float* __restrict__ row, int t, bool valid) {
  Raw8 r;
  r.x = make_float4(0.f, 0.f, 0.f, 0.f);
  r.y = r.x;
  if (valid) {
    r.x = ldg4(row + 4 * t);
    r.y = ldg4(row + 16 + 4 * t);
  }
  return r;
}
__device__ __forceinline__ void split_row(Frag8& f, const Raw8& r, float scale) {
  float v[8] = {r.x.x, r.x.y, r.x.z, r.x.w, r.y.x, r.y.y, r.y.z, r.y.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (scale != 1.f) v[i] *= scale;
    split_tf32(v[i], f.hi[i], f.lo[i]);
  }
}
// rows 8kt + 2t and 8kt + 2t + 1, columns coff..coff+3 (column pattern)
__device__ __forceinline__ Raw8 load_cols_raw(const float* __restrict__ base, int rs, int kt, int t, int coff,
                                              int kb) {
  const int ja = kt * 8 + 2 * t, jb = ja + 1;
  Raw8 r;
  r.x = make_float4(0.f, 0.f, 0.f, 0.f);
  r.y = r.x;
  if (ja < kb) r.x = ldg4(base + ja * rs + coff);
  if (jb < kb) r.y = ldg4(base + jb * rs + coff);
  return r;
}
// d = A(rows g, g+8) . B(row 8nt + g)^T over the 32 columns (row-pattern fragments on both sides).  The three products
// of the compensated scheme run as three independent accumulation chains of 4 (a warp issues in order: twelve dependent
// MMAs would expose the instruction's latency twelve times); small terms are summed first.
__device__ __forceinline__ void mma_rows(float (&d)[4], const Frag8& a0, const Frag8& a1, const Frag8& b) {
  float lh[4] = {0.f, 0.f, 0.f, 0.f}, hl[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    mma_tf32(lh, a0.lo[2 * kk], a1.lo[2 * kk], a0.lo[2 * kk + 1], a1.lo[2 * kk + 1], b.hi[2 * kk], b.hi[2 * kk + 1]);
    mma_tf32(hl, a0.hi[2 * kk], a1.hi[2 * kk], a0.hi[2 * kk + 1], a1.hi[2 * kk + 1], b.lo[2 * kk], b.lo[2 * kk + 1]);
    mma_tf32(d, a0.hi[2 * kk], a1.hi[2 * kk], a0.hi[2 * kk + 1], a1.hi[2 * kk + 1], b.hi[2 * kk], b.hi[2 * kk + 1]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i] += lh[i] + hl[i];
}
// acc[ntd] += P(rows g, g+8; keys 8kt + 2t, +1 = p[0..3] in accumulator layout) . B(rows 8kt + 2t, +1; column pattern)
__device__ __forceinline__ void mma_cols(float (&acc)[4][4], const float (&p)[4], const Raw8& r, float scale) {
  uint32_t ah[4], al[4];
  split_tf32(p[0], ah[0], al[0]);   // (g,   k = t)
  split_tf32(p[2], ah[1], al[1]);   // (g+8, k = t)
  split_tf32(p[1], ah[2], al[2]);   // (g,   k = t+4)
  split_tf32(p[3], ah[3], al[3]);   // (g+8, k = t+4)
  float b0[4] = {r.x.x, r.x.y, r.x.z, r.x.w}, b1[4] = {r.y.x, r.y.y, r.y.z, r.y.w};
  uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
  for (int ntd = 0; ntd < 4; ++ntd) {
    if (scale != 1.f) {
      b0[ntd] *= scale;
      b1[ntd] *= scale;
    }
    split_tf32(b0[ntd], bh0[ntd], bl0[ntd]);
    split_tf32(b1[ntd], bh1[ntd], bl1[ntd]);
  }
  // product-major order: four independent accumulators between two instructions on the same one
#pragma unroll
  for (int ntd = 0; ntd < 4; ++ntd) mma_tf32(acc[ntd], al[0], al[1], al[2], al[3], bh0[ntd], bh1[ntd]);
#pragma unroll
  for (int ntd = 0; ntd < 4; ++ntd) mma_tf32(acc[ntd], ah[0], ah[1], ah[2], ah[3], bl0[ntd], bl1[ntd]);
#pragma unroll
  for (int ntd = 0; ntd < 4; ++ntd) mma_tf32(acc[ntd], ah[0], ah[1], ah[2], ah[3], bh0[ntd], bh1[ntd]);
}
// rows g and g+8 of an output tile: columns 4t..4t+3 from acc[.][0|2], 16+4t.. from acc[.][1|3]
__device__ __forceinline__ void store_rows(float* __restrict__ base, int rs, int ja, int jb, int kb, int t,
                                           const float (&acc)[4][4], float scale) {
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = scale != 1.f ? acc[i][j] * scale : acc[i][j];
  if (ja < kb) {
    *reinterpret_cast<float4*>(base + ja * rs + 4 * t) = make_float4(o[0][0], o[1][0], o[2][0], o[3][0]);
    *reinterpret_cast<float4*>(base + ja * rs + 16 + 4 * t) = make_float4(o[0][1], o[1][1], o[2][1], o[3][1]);
  }
  if (jb < kb) {
    *reinterpret_cast<float4*>(base + jb * rs + 4 * t) = make_float4(o[0][2], o[1][2], o[2][2], o[3][2]);
    *reinterpret_cast<float4*>(base + jb * rs + 16 + 4 * t) = make_float4(o[0][3], o[1][3], o[2][3], o[3][3]);
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// every token row of this head (128 B each) is requested from DRAM into L2 before the first dependent load: the loads
// that follow find it on its way instead of paying one DRAM round trip per phase
__device__ __forceinline__ void prefetch_rows(const float* __restrict__ base, int rs, int kb) {
  for (int j = threadIdx.x & 31; j < kb; j += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + j * rs));
}

struct AmCtx {
  long long node, r0;
  int rs;   // float stride between the node's tokens (32-bit: am_ok bounds 48 * N * ld)
  int h, kb, NT, g, t, coff;
};
__device__ __forceinline__ bool am_ctx(const AttArgs& a, AmCtx& c) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  c.node = blockIdx.x;
  c.h = blockIdx.y * AM_WARPS + warp;
  if (c.h >= a.n_head) return false;
  const int b = (int)a.batch[c.node];
  const int node0 = a.graph_ptr[b];
  const int n = a.graph_ptr[b + 1] - node0;
  c.kb = a.masked ? (n < a.kslots ? n : a.kslots) : a.kslots;
  c.r0 = (a.row_ptr[b] + (c.node - node0)) * a.ld + c.h * 32;   // float offset of token 0, this head
  c.rs = n * (int)a.ld;
  c.NT = (c.kb + 7) >> 3;
  c.g = lane >> 2;
  c.t = lane & 3;
  c.coff = ((c.g & 1) << 4) + ((c.g >> 1) << 2);
  return true;
}

// scores (in units of ln 2: the query scale carries log2 e) of rows (ja, jb) against every key -> probabilities in s
// (accumulator layout); returns row max and 1 / row sum
__device__ __forceinline__ void softmax_rows(float (&s)[AM_NT][4], int NT, int kb, int t, float& ma, float& mb,
                                             float& la, float& lb) {
  ma = -INFINITY;
  mb = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < AM_NT; ++nt) {
    if (nt < NT) {
      const int c0 = nt * 8 + 2 * t;
      if (c0 >= kb) s[nt][0] = s[nt][2] = -INFINITY;
      if (c0 + 1 >= kb) s[nt][1] = s[nt][3] = -INFINITY;
      ma = fmaxf(ma, fmaxf(s[nt][0], s[nt][1]));
      mb = fmaxf(mb, fmaxf(s[nt][2], s[nt][3]));
    }
  }
  ma = quad_max(ma);
  mb = quad_max(mb);
  la = 0.f;
  lb = 0.f;
#pragma unroll
  for (int nt = 0; nt < AM_NT; ++nt) {
    if (nt < NT) {
      s[nt][0] = exp2f(s[nt][0] - ma);
      s[nt][1] = exp2f(s[nt][1] - ma);
      s[nt][2] = exp2f(s[nt][2] - mb);
      s[nt][3] = exp2f(s[nt][3] - mb);
      la += s[nt][0] + s[nt][1];
      lb += s[nt][2] + s[nt][3];
    }
  }
  la = __frcp_rn(quad_sum(la));   // returned as 1 / row sum
  lb = __frcp_rn(quad_sum(lb));
#pragma unroll
  for (int nt = 0; nt < AM_NT; ++nt) {
    if (nt < NT) {
      s[nt][0] *= la;
      s[nt][1] *= la;
      s[nt][2] *= lb;
      s[nt][3] *= lb;
    }
  }
}

// row-pattern raw loads of every key tile (rows 8nt + g)
__device__ __forceinline__ void rows_load(Raw8 (&raw)[AM_NT], const float* __restrict__ bp, const AmCtx& c) {
#pragma unroll
  for (int nt = 0; nt < AM_NT; ++nt)
    if (nt < c.NT) raw[nt] = load_raw(bp + (nt * 8 + c.g) * c.rs, c.t, nt * 8 + c.g < c.kb);
}
// s[nt] = A(rows ja, jb) . B(rows 8nt + g)^T for every key tile
__device__ __forceinline__ void rows_product(float (&s)[AM_NT][4], const Frag8& a0, const Frag8& a1,
                                             const Raw8 (&raw)[AM_NT], int NT, float bscale) {
#pragma unroll
  for (int nt = 0; nt < AM_NT; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    if (nt < NT) {
      Frag8 f;
      split_row(f, raw[nt], bscale);
      mma_rows(s[nt], a0, a1, f);
    }
  }
}
__device__ __forceinline__ void cols_load(Raw8 (&raw)[AM_NT], const float* __restrict__ bp, const AmCtx& c) {
#pragma unroll
  for (int kt = 0; kt < AM_NT; ++kt)
    if (kt < c.NT) raw[kt] = load_cols_raw(bp, c.rs, kt, c.t, c.coff, c.kb);
}
__device__ __forceinline__ void cols_product(float (&acc)[4][4], const float (&p)[AM_NT][4], const Raw8 (&raw)[AM_NT],
                                             int NT, float bscale) {
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < AM_NT; ++kt)
    if (kt < NT) mma_cols(acc, p[kt], raw[kt], bscale);
}

template <bool DROP>   // DROP: training-mode attention dropout (the reference's quirk); compiled out otherwise
__global__ void __launch_bounds__(32 * AM_WARPS, 3) attention_mma_fwd_kernel(const AttArgs a) {
  AmCtx c;
  if (!am_ctx(a, c)) return;
  const float* qp = a.q + c.r0;
  const float* kp = a.k + c.r0;
  const float* vp = a.v + c.r0;
  float* op = a.o + c.r0;
  const int g = c.g, t = c.t, kb = c.kb, NT = c.NT;
  const float rTs = 1.4426950408889634f / a.inv_temp_div;   // scores in units of ln 2 (softmax by exp2)
  prefetch_rows(qp, c.rs, kb);
  prefetch_rows(kp, c.rs, kb);
  prefetch_rows(vp, c.rs, kb);
  for (int mt = 0; mt * 16 < kb; ++mt) {
    const int ja = mt * 16 + g, jb = ja + 8;
    float s[AM_NT][4];
    {
      const Raw8 ra = load_raw(qp + ja * c.rs, t, ja < kb), rb = load_raw(qp + jb * c.rs, t, jb < kb);
      Raw8 kraw[AM_NT];
      rows_load(kraw, kp, c);
      Frag8 qa, qb;
      split_row(qa, ra, rTs);
      split_row(qb, rb, rTs);
      rows_product(s, qa, qb, kraw, NT, 1.f);
    }
    Raw8 vraw[AM_NT];
    cols_load(vraw, vp, c);   // in flight during the softmax
    float ma, mb, la, lb;
    softmax_rows(s, NT, kb, t, ma, mb, la, lb);
    if (DROP) {
#pragma unroll
      for (int nt = 0; nt < AM_NT; ++nt) {
        if (nt < NT) {
          const int c0 = nt * 8 + 2 * t;
          s[nt][0] *= att_keep_scale(a.seed, c.node, c.h, ja, c0, a.drop_p);
          s[nt][1] *= att_keep_scale(a.seed, c.node, c.h, ja, c0 + 1, a.drop_p);
          s[nt][2] *= att_keep_scale(a.seed, c.node, c.h, jb, c0, a.drop_p);
          s[nt][3] *= att_keep_scale(a.seed, c.node, c.h, jb, c0 + 1, a.drop_p);
        }
      }
    }
    float o[4][4];
    cols_product(o, s, vraw, NT, 1.f);
    store_rows(op, c.rs, ja, jb, kb, t, o, 1.f);
  }
}

#define AM_TR_ROWS 48      // query rows kept per warp (3 row tiles)
#define AM_TR_LD 44        // row stride of the transposition buffers: 2*44 = 24 (mod 32) -> phase B's reads hit 32 banks
#define AM_BWD_SMEM (AM_WARPS * 2 * AM_TR_ROWS * AM_TR_LD * (int)sizeof(float))

// Backward.  Phase A (rows = queries): S = (Q/T) K^T and dP = dO V^T -> P, D = sum_j2 dP P, dS = P (dP - D), dQ = dS K / T;
// Pdrop and dS are also written to shared memory [query][key].  Phase B (rows = keys) reads them back TRANSPOSED as the
// A fragments of dV = Pdrop^T dO and dK = dS^T (Q/T): no recomputation of the scores (the first version recomputed
// S^T = K Q^T and dP^T = V dO^T with the operands swapped: two of seven contractions, their loads, splits and exp2).
template <bool DROP>
__global__ void __launch_bounds__(32 * AM_WARPS, 3) attention_mma_bwd_kernel(const AttArgs a) {
  extern __shared__ __align__(16) float am_smem[];
  AmCtx c;
  if (!am_ctx(a, c)) return;
  const int warp = threadIdx.x >> 5;
  float* Pt = am_smem + (size_t)warp * 2 * AM_TR_ROWS * AM_TR_LD;   // Pdrop[j1][j2]
  float* St = Pt + AM_TR_ROWS * AM_TR_LD;                            // dS[j1][j2]
  const float* qp = a.q + c.r0;
  const float* kp = a.k + c.r0;
  const float* vp = a.v + c.r0;
  const float* gp = a.go + c.r0;
  const int g = c.g, t = c.t, kb = c.kb, NT = c.NT;
  const float rT = 1.f / a.inv_temp_div;
  const float rTs = 1.4426950408889634f / a.inv_temp_div;   // scores in units of ln 2 (softmax by exp2)
  prefetch_rows(qp, c.rs, kb);
  prefetch_rows(gp, c.rs, kb);
  prefetch_rows(kp, c.rs, kb);
  prefetch_rows(vp, c.rs, kb);
  // ---- phase A
  for (int mt = 0; mt * 16 < kb; ++mt) {
    const int ja = mt * 16 + g, jb = ja + 8;
    float s[AM_NT][4], dp[AM_NT][4];
    {
      const Raw8 r0 = load_raw(qp + ja * c.rs, t, ja < kb), r1 = load_raw(qp + jb * c.rs, t, jb < kb);
      const Raw8 r2 = load_raw(gp + ja * c.rs, t, ja < kb), r3 = load_raw(gp + jb * c.rs, t, jb < kb);
      Raw8 raw[AM_NT];
      rows_load(raw, kp, c);
      Frag8 fa, fb;
      split_row(fa, r0, rTs);
      split_row(fb, r1, rTs);
      rows_product(s, fa, fb, raw, NT, 1.f);
      rows_load(raw, vp, c);
      split_row(fa, r2, 1.f);
      split_row(fb, r3, 1.f);
      rows_product(dp, fa, fb, raw, NT, 1.f);
    }
    Raw8 kraw[AM_NT];
    cols_load(kraw, kp, c);
    float ma, mb, la, lb;
    softmax_rows(s, NT, kb, t, ma, mb, la, lb);
    const bool va = ja < kb, vb = jb < kb;   // query rows beyond the token set contribute nothing to dK / dV
    float da = 0.f, db = 0.f;
#pragma unroll
    for (int nt = 0; nt < AM_NT; ++nt) {
      if (nt < NT) {
        const int c0 = nt * 8 + 2 * t;
        float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
        if (DROP) {
          k0 = att_keep_scale(a.seed, c.node, c.h, ja, c0, a.drop_p);
          k1 = att_keep_scale(a.seed, c.node, c.h, ja, c0 + 1, a.drop_p);
          k2 = att_keep_scale(a.seed, c.node, c.h, jb, c0, a.drop_p);
          k3 = att_keep_scale(a.seed, c.node, c.h, jb, c0 + 1, a.drop_p);
          dp[nt][0] *= k0; dp[nt][1] *= k1; dp[nt][2] *= k2; dp[nt][3] *= k3;
        }
        *reinterpret_cast<float2*>(Pt + ja * AM_TR_LD + c0) = va ? make_float2(s[nt][0] * k0, s[nt][1] * k1) : make_float2(0.f, 0.f);
        *reinterpret_cast<float2*>(Pt + jb * AM_TR_LD + c0) = vb ? make_float2(s[nt][2] * k2, s[nt][3] * k3) : make_float2(0.f, 0.f);
        da = fmaf(dp[nt][0], s[nt][0], da);
        da = fmaf(dp[nt][1], s[nt][1], da);
        db = fmaf(dp[nt][2], s[nt][2], db);
        db = fmaf(dp[nt][3], s[nt][3], db);
      }
    }
    da = quad_sum(da);
    db = quad_sum(db);
#pragma unroll
    for (int nt = 0; nt < AM_NT; ++nt) {
      if (nt < NT) {
        const int c0 = nt * 8 + 2 * t;
        s[nt][0] *= dp[nt][0] - da;
        s[nt][1] *= dp[nt][1] - da;
        s[nt][2] *= dp[nt][2] - db;
        s[nt][3] *= dp[nt][3] - db;
        *reinterpret_cast<float2*>(St + ja * AM_TR_LD + c0) = va ? make_float2(s[nt][0], s[nt][1]) : make_float2(0.f, 0.f);
        *reinterpret_cast<float2*>(St + jb * AM_TR_LD + c0) = vb ? make_float2(s[nt][2], s[nt][3]) : make_float2(0.f, 0.f);
      }
    }
    float dq[4][4];
    cols_product(dq, s, kraw, NT, 1.f);
    store_rows(a.gq + c.r0, c.rs, ja, jb, kb, t, dq, rT);
  }
  __syncwarp();
  // ---- phase B: rows = keys; the contraction runs over the queries 8kt + 2t, 8kt + 2t + 1 (same permutation as the
  // column-pattern B operand)
  for (int mt = 0; mt * 16 < kb; ++mt) {
    const int ja = mt * 16 + g, jb = ja + 8;   // key rows of this lane
    Raw8 graw[AM_NT], qraw[AM_NT];
    cols_load(graw, gp, c);
    cols_load(qraw, qp, c);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < AM_NT; ++kt) {
      if (kt < NT) {
        const float* r0 = Pt + (kt * 8 + 2 * t) * AM_TR_LD;
        const float p[4] = {ja < kb ? r0[ja] : 0.f, ja < kb ? r0[AM_TR_LD + ja] : 0.f, jb < kb ? r0[jb] : 0.f,
                            jb < kb ? r0[AM_TR_LD + jb] : 0.f};   // key rows beyond the token set: not stored, not read
        mma_cols(acc, p, graw[kt], 1.f);
      }
    }
    store_rows(a.gv + c.r0, c.rs, ja, jb, kb, t, acc, 1.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < AM_NT; ++kt) {
      if (kt < NT) {
        const float* r0 = St + (kt * 8 + 2 * t) * AM_TR_LD;
        const float p[4] = {ja < kb ? r0[ja] : 0.f, ja < kb ? r0[AM_TR_LD + ja] : 0.f, jb < kb ? r0[jb] : 0.f,
                            jb < kb ? r0[AM_TR_LD + jb] : 0.f};   // key rows beyond the token set: not stored, not read
        mma_cols(acc, p, qraw[kt], rT);
      }
    }
    store_rows(a.gk + c.r0, c.rs, ja, jb, kb, t, acc, 1.f);
  }
}

int g_att_mma = 1;

bool am_ok(const AttArgs& a, int kmax, bool bwd) {
  auto al = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
  if (!g_att_mma || a.dk != 32 || kmax > AM_KMAX || (a.ld & 3) != 0) return false;
  if (48ll * a.N * a.ld >= (1ll << 31)) return false;   // token offsets inside a node are 32-bit
  if (!(al(a.q) && al(a.k) && al(a.v))) return false;
  return bwd ? (al(a.go) && al(a.gq) && al(a.gk) && al(a.gv)) : al(a.o);
}
}  // namespace

extern "C" int sb_set_attention_mma(int32_t enable) {
  const int old = g_att_mma;
  g_att_mma = enable ? 1 : 0;
  return old;
}

int sb_attention_mma_fwd_launch(const AttArgs& a, int kmax, cudaStream_t st) {
  if (!am_ok(a, kmax, false)) return SB_ERR_UNSUPPORTED;
  dim3 grid((unsigned)a.N, (unsigned)((a.n_head + AM_WARPS - 1) / AM_WARPS));
  if (a.drop_p > 0.f) attention_mma_fwd_kernel<true><<<grid, 32 * AM_WARPS, 0, st>>>(a);
  else attention_mma_fwd_kernel<false><<<grid, 32 * AM_WARPS, 0, st>>>(a);
  SB_CHECK_LAUNCH("sb_attention_fwd(mma)");
  return SB_OK;
}
int sb_attention_mma_bwd_launch(const AttArgs& a, int kmax, cudaStream_t st) {
  if (!am_ok(a, kmax, true)) return SB_ERR_UNSUPPORTED;
  dim3 grid((unsigned)a.N, (unsigned)((a.n_head + AM_WARPS - 1) / AM_WARPS));
  static bool configured = false;
  if (!configured) {
    SB_CUDA(cudaFuncSetAttribute(attention_mma_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AM_BWD_SMEM));
    SB_CUDA(cudaFuncSetAttribute(attention_mma_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AM_BWD_SMEM));
    configured = true;
  }
  if (a.drop_p > 0.f) attention_mma_bwd_kernel<true><<<grid, 32 * AM_WARPS, AM_BWD_SMEM, st>>>(a);
  else attention_mma_bwd_kernel<false><<<grid, 32 * AM_WARPS, AM_BWD_SMEM, st>>>(a);
  SB_CHECK_LAUNCH("sb_attention_bwd(mma)");
  return SB_OK;
}

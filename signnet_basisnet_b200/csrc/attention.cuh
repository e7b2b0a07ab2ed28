// Shared between transformer.cu (generic attention kernels) and attention_fast.cu (register-resident fast path).
#pragma once
#include "common.cuh"

struct AttArgs {
  const float* q;
  const float* k;
  const float* v;
  long long ld;      // row stride of q/k/v/o (and their gradients)
  const int64_t* batch;
  const int32_t* graph_ptr;
  const int64_t* row_ptr;
  long long N;
  int kslots, masked, n_head, dk;
  float inv_temp_div;  // temperature (sqrt(dk)); q is divided by it
  float drop_p;
  unsigned long long seed;
  float* o;
  // backward only
  const float* go;
  float* gq;
  float* gk;
  float* gv;
};

__device__ __forceinline__ float att_keep_scale(unsigned long long seed, long long node, int h, int j1, int j2,
                                                float p) {
  // counter-based generator -> uniform in [0,1); the same mask is regenerated in the backward.  The 64-bit part depends
  // on (seed, node, head, query) only - loop-invariant along a score row, the compiler hoists it - and the per-key part
  // is one 32-bit avalanche mixer (lowbias32): ~8 integer instructions per attention probability instead of the ~30 of
  // the three 64-bit multiplies of a full splitmix64 round (25 % of the attention kernels' time in training mode).
  const unsigned long long key = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(((node * 64 + h) * 4096 + j1) + 1);
  unsigned int x = (unsigned int)(key ^ (key >> 29)) + 0x85EBCA6Bu * (unsigned int)(j2 + 1);
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  const float u = (float)(x >> 8) * (1.0f / 16777216.0f);
  return (u >= p) ? 1.0f / (1.0f - p) : 0.0f;
}


// fast path: k_b <= 40 tokens, d_k <= 32 (covers every shipped configuration); SB_ERR_UNSUPPORTED otherwise
int sb_attention_fast_fwd_launch(const AttArgs& a, int kmax, cudaStream_t st);
int sb_attention_fast_bwd_launch(const AttArgs& a, int kmax, cudaStream_t st);
// tensor-core path (attention_mma.cu): d_k == 32, k_b <= 40, 16-byte aligned rows; SB_ERR_UNSUPPORTED otherwise
int sb_attention_mma_fwd_launch(const AttArgs& a, int kmax, cudaStream_t st);
int sb_attention_mma_bwd_launch(const AttArgs& a, int kmax, cudaStream_t st);

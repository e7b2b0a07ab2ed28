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
  // counter-based hash (splitmix64) -> uniform in [0,1); the same mask is regenerated in the backward
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(((node * 64 + h) * 4096 + j1) * 4096 + j2 + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
  return (u >= p) ? 1.0f / (1.0f - p) : 0.0f;
}


// fast path: k_b <= 40 tokens, d_k <= 32 (covers every shipped configuration); SB_ERR_UNSUPPORTED otherwise
int sb_attention_fast_fwd_launch(const AttArgs& a, int kmax, cudaStream_t st);
int sb_attention_fast_bwd_launch(const AttArgs& a, int kmax, cudaStream_t st);
// tensor-core path (attention_mma.cu): d_k == 32, k_b <= 40, 16-byte aligned rows; SB_ERR_UNSUPPORTED otherwise
int sb_attention_mma_fwd_launch(const AttArgs& a, int kmax, cudaStream_t st);
int sb_attention_mma_bwd_launch(const AttArgs& a, int kmax, cudaStream_t st);

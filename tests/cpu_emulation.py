"""Serial CPU emulation of CUDA kernels that have no inter-thread communication (no shuffles other than a lane-0
broadcast, no barriers, shared memory or atomics): their SOURCE TEXT is compiled with g++ under a tiny prelude that defines
threadIdx / __ldg / float4 ... and launched thread by thread.  TEST INFRASTRUCTURE ONLY: this is how kernel files written
without GPU access get their indexing and arithmetic checked; nothing in signnet_basisnet_b200/ can reach it."""
import ctypes
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRELUDE = r"""
#include <cmath>
#include <cstdint>
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
struct uint3_ { unsigned x, y, z; };
static uint3_ threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
static inline void stg4_stream(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// broadcast from lane 0 only (the one pattern canonical_sign_kernel uses): lanes run in order 0..31, so lane 0 has
// already deposited its value when the other lanes of the warp ask for it
static float shfl_lane0_;
static inline float __shfl_sync(unsigned, float v, int src_lane) {
  if (src_lane != 0) __builtin_trap();
  if ((threadIdx.x & 31) == 0) shfl_lane0_ = v;
  return shfl_lane0_;
}
#define LAUNCH(kernel, grid, ...)                                                       \
  for (unsigned b_ = 0; b_ < (unsigned)(grid); ++b_)                                    \
    for (unsigned t_ = 0; t_ < 256; ++t_) {                                             \
      blockIdx = {b_, 0, 0}; threadIdx = {t_, 0, 0}; blockDim = {256, 1, 1}; gridDim = {(unsigned)(grid), 1, 1};  \
      kernel(__VA_ARGS__);                                                              \
    }
"""


def functions(src, pattern):
    out = []
    for m in re.finditer(pattern, src):
        k = src.index("{", m.start())
        depth = 0
        while True:
            depth += {"{": 1, "}": -1}.get(src[k], 0)
            if depth == 0:
                break
            k += 1
        out.append(src[m.start():k + 1])
    return out



def build(tmp_dir, cu_file, patterns, wrappers, defines=""):
    """Extract the functions matching `patterns` from csrc/<cu_file>, append extern "C" `wrappers` (which use LAUNCH)
    and return the loaded shared library."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = open(os.path.join(ROOT, "signnet_basisnet_b200", "csrc", cu_file)).read()
    parts = [f for pat in patterns for f in functions(src, pat)]
    cpp, so = os.path.join(tmp_dir, "emu.cpp"), os.path.join(tmp_dir, "libemu.so")
    open(cpp, "w").write(PRELUDE + defines + "\n" + "\n".join(parts) + "\n" + wrappers)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp], check=True)
    return ctypes.CDLL(so), len(parts)


def stable_csr(keys, other, N):
    """What sb_build_csr produces: rows by `keys`, edge-id order inside a row; (ptr, neighbour, edge id) as int32."""
    import torch

    order = torch.sort(keys, stable=True).indices
    ptr = torch.zeros(N + 1, dtype=torch.int32)
    ptr[1:] = torch.cumsum(torch.bincount(keys, minlength=N), 0).to(torch.int32)
    return ptr, other[order].to(torch.int32).contiguous(), order.to(torch.int32).contiguous()

// The whole phi stack of SignNet (L masked-GIN layers, both sign passes) behind TWO C-ABI calls.
//
// Reference: GNN3d.forward (Alchemy/sign_net/sign_net.py:28-44) = L x { MaskedGINConv (model_utils/masked_layers.py:74-84:
// GINConv aggregate -> MaskedMLP :54-64) -> MaskedBN -> ReLU -> residual }, invoked for +v and -v (sign_net.py:113).
//
// Why this exists: the per-kernel entry points of this library were driven from Python, ~20 C-ABI calls per layer
// (forward 6, backward 14) plus one torch allocation per intermediate; at 128 graphs per GPU (BASELINE.json configs[3]
// sharded 8 ways) the GPU needs ~5 ms for a step and the Python side ~10 ms to issue it (scripts/host_probe.py).  Here
// the layer loop is host C++: the caller hands over a table of device pointers (activations to save, parameters,
// BatchNorm buffers, scratch) and the launch sequence - exactly the one signnet_basisnet_b200/phi.py documents - is
// enqueued back to back on the caller's stream.  No allocation, no synchronisation, nothing retained.
#include <cuda_runtime.h>
#include "common.cuh"
#include "../../include/signnet_b200.h"

namespace {

template <typename T>
inline T* P(int64_t v) { return reinterpret_cast<T*>(static_cast<uintptr_t>(v)); }
inline int pad4i(int d) { return (d + 3) / 4 * 4; }

struct Slots {
  const int32_t *graph_ptr, *unit_ptr, *unit_desc, *in_ptr, *in_src, *out_ptr, *out_dst;
  const uint32_t *in_pack, *out_pack;
  const int64_t* row_ptr;
  int64_t R;
  int B, k, masked, tile_rows, generic;
};

Slots read_slots(const int64_t* p, const int64_t* n) {
  Slots s;
  s.graph_ptr = P<const int32_t>(p[0]); s.unit_ptr = P<const int32_t>(p[1]); s.unit_desc = P<const int32_t>(p[2]);
  s.in_pack = P<const uint32_t>(p[3]); s.out_pack = P<const uint32_t>(p[4]); s.row_ptr = P<const int64_t>(p[5]);
  s.in_ptr = P<const int32_t>(p[6]); s.in_src = P<const int32_t>(p[7]);
  s.out_ptr = P<const int32_t>(p[8]); s.out_dst = P<const int32_t>(p[9]);
  s.R = n[0]; s.B = (int)n[1]; s.k = (int)n[2]; s.masked = (int)n[3]; s.tile_rows = (int)n[4]; s.generic = (int)n[5];
  return s;
}

int agg(const Slots& s, const float* x, float* out, const float* res, const float* dotx, double* dot_out,
        const float* eps, int S, int ld, bool transpose, void* st) {
  const int generic = (s.generic || (ld % 4) != 0) ? 1 : 0;
  return sb_gin_agg(x, out, res, dotx, dot_out, eps, s.graph_ptr, s.unit_ptr, s.unit_desc,
                    transpose ? s.out_pack : s.in_pack, s.row_ptr, transpose ? s.out_ptr : s.in_ptr,
                    transpose ? s.out_dst : s.in_src, s.R, s.B, s.k, s.masked, S, ld, s.tile_rows > 1 ? s.tile_rows : 1,
                    generic, st);
}

// d act(BN(y)) -> d y :  reduction (reads gout, y) -> coefficients -> apply (recomputes the ReLU mask), dz may alias gout
// `stats`: this BatchNorm's own zero-initialised fp64 [S,2,C] region of the caller's arena (one fill per backward instead of a
// memset per BatchNorm: every launch-queue entry counts)
int bn_backward(const float* gout, const float* y, const float* a, const float* c, const double* mr, const float* gamma,
                int64_t ld, int64_t R, int S, int C, int training, float* dz, float* dgamma, float* dbeta, double* stats,
                cudaStream_t st) {
  const int rc = sb_bn_bwd_reduce(gout, y, a, c, mr, nullptr, ld, R, S, C, 1, stats, st);
  if (rc) return rc;
  // coefficients are derived inside the apply kernel (sb_bn_apply_bwd: one launch instead of two)
  return sb_bn_apply_bwd(gout, y, stats, mr, a, c, gamma, R, training, dz, dgamma, dbeta, ld, R, S, C, st);
}

}  // namespace

#define PHI_FWD_COLS 25
#define PHI_BWD_COLS 23

// layer_ptrs[l] = { X_in, A, H, Y, X_out,  W0, g0, b0, W1, b1|0, eps, g1, bb1,  rm0, rv0, rm1, rv1,  st0|0, st1|0,
//                   a0, c0, mr0, a1, c1, mr1 }   (st*: fp64 [S,2,C] zero-initialised by the caller; 0 in eval mode)
// dims[l] = { d_in, h, d, ld_in }   (ld_in = 1 for the first layer's [S,R] input, else the padded row stride)
// slot_ptrs[2][10], slot_ints[2][6]: row 0 = layout for padded rows, row 1 = layout used when ld_in % 4 != 0 (layer 0)
extern "C" int sb_phi_stack_fwd(const int64_t* layer_ptrs, const int32_t* dims, int32_t L, const int64_t* slot_ptrs,
                                const int64_t* slot_ints, int32_t S, int32_t training, float momentum, float bn_eps,
                                void* stream) {
  SB_CHECK_ARG(layer_ptrs && dims && slot_ptrs && slot_ints && L >= 1 && S >= 1, "sb_phi_stack_fwd: bad arguments");
  const Slots main_sl = read_slots(slot_ptrs, slot_ints), in_sl = read_slots(slot_ptrs + 10, slot_ints + 6);
  const int64_t R = main_sl.R;
  for (int l = 0; l < L; ++l) {
    const int64_t* p = layer_ptrs + (size_t)l * PHI_FWD_COLS;
    const int d_in = dims[4 * l], h = dims[4 * l + 1], d = dims[4 * l + 2], ld_in = dims[4 * l + 3];
    const int ldh = pad4i(h), ldd = pad4i(d);
    const Slots& sl = (l == 0 && (ld_in % 4) != 0) ? in_sl : main_sl;
    const float* X = P<const float>(p[0]);
    float *A = P<float>(p[1]), *H = P<float>(p[2]), *Y = P<float>(p[3]), *Xn = P<float>(p[4]);
    // aggregate + first Linear: one fused launch where the shape allows it (gin_lin_fused.cu: A stays on the SM on its
    // way into the contraction), else the two kernels - identical results either way
    int rc = sb_gin_linear_fused_fwd(X, A, H, P<double>(p[17]), P<const float>(p[10]), P<const float>(p[5]), d_in, 1, d_in,
                                     h, ldh, sl.unit_ptr, sl.unit_desc, sl.in_pack, sl.in_ptr, sl.in_src, sl.R, sl.B, S,
                                     ld_in, sl.tile_rows, sl.generic, stream);
    if (rc == SB_ERR_UNSUPPORTED) {
      rc = agg(sl, X, A, nullptr, nullptr, nullptr, P<const float>(p[10]), S, ld_in, false, stream);
      if (rc) return rc;
      rc = sb_linear_fwd(A, ld_in, P<const float>(p[5]), d_in, 1, nullptr, H, ldh, R, S, d_in, h, 0, nullptr, nullptr, 0,
                         P<double>(p[17]), 0, stream);
    }
    if (rc) return rc;
    rc = sb_bn_finalize(P<const double>(p[17]), R, S, h, P<const float>(p[6]), P<const float>(p[7]), P<float>(p[13]),
                        P<float>(p[14]), momentum, bn_eps, training, P<float>(p[19]), P<float>(p[20]), P<double>(p[21]),
                        stream);
    if (rc) return rc;
    rc = sb_linear_fwd(H, ldh, P<const float>(p[8]), h, 1, P<const float>(p[9]), Y, ldd, R, S, h, d, 2,
                       P<const float>(p[19]), P<const float>(p[20]), 0, P<double>(p[18]), 0, stream);
    if (rc) return rc;
    // outer BatchNorm: finalize + relu(a Y + c) + residual in one launch
    rc = sb_bn_apply_fwd(Y, P<const double>(p[18]), R, S, d, P<const float>(p[11]), P<const float>(p[12]), P<float>(p[15]),
                         P<float>(p[16]), momentum, bn_eps, training, 1, l > 0 ? X : nullptr, Xn, ldd, R, P<float>(p[22]),
                         P<float>(p[23]), P<double>(p[24]), stream);
    if (rc) return rc;
  }
  return SB_OK;
}

// layer_ptrs[l] = { X, A, H, Y,  a0, c0, mr0, a1, c1, mr1,  W0, g0, W1, eps, g1,
//                   gW0, gg0, gb0, gW1, gb1|0, deps (fp64 scalar, zero-initialised), gg1, gbb1 }
// scratch = { G (in: dL/dX_L; updated in place down the residual stream), dY, dH, dA, out0 (layer-0 aggregate sink),
//             stats arena fp64 [2L][S,2,Cmax] ZERO-INITIALISED by the caller (region 2l = outer, 2l+1 = inner BatchNorm of
//             layer l), region stride in doubles (an integer, not a pointer), wgrad workspace
//             (sb_linear_wgrad_workspace_floats) }
extern "C" int sb_phi_stack_bwd(const int64_t* layer_ptrs, const int32_t* dims, int32_t L, const int64_t* slot_ptrs,
                                const int64_t* slot_ints, const int64_t* scratch, int32_t S, int32_t training,
                                void* stream) {
  SB_CHECK_ARG(layer_ptrs && dims && slot_ptrs && slot_ints && scratch && L >= 1 && S >= 1,
               "sb_phi_stack_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const Slots main_sl = read_slots(slot_ptrs, slot_ints), in_sl = read_slots(slot_ptrs + 10, slot_ints + 6);
  const int64_t R = main_sl.R;
  float *G = P<float>(scratch[0]), *dY = P<float>(scratch[1]), *dH = P<float>(scratch[2]), *dA = P<float>(scratch[3]),
        *out0 = P<float>(scratch[4]);
  double* stats = P<double>(scratch[5]);
  const int64_t stats_stride = scratch[6];
  float* ws = P<float>(scratch[7]);
  for (int l = L - 1; l >= 0; --l) {
    const int64_t* p = layer_ptrs + (size_t)l * PHI_BWD_COLS;
    const int d_in = dims[4 * l], h = dims[4 * l + 1], d = dims[4 * l + 2], ld_in = dims[4 * l + 3];
    const int ldh = pad4i(h), ldd = pad4i(d);
    const Slots& sl = (l == 0 && (ld_in % 4) != 0) ? in_sl : main_sl;
    const float *X = P<const float>(p[0]), *A = P<const float>(p[1]), *H = P<const float>(p[2]), *Y = P<const float>(p[3]);
    const float *a0 = P<const float>(p[4]), *c0 = P<const float>(p[5]), *a1 = P<const float>(p[7]), *c1 = P<const float>(p[8]);
    const double *mr0 = P<const double>(p[6]), *mr1 = P<const double>(p[9]);
    // outer BN + ReLU
    int rc = bn_backward(G, Y, a1, c1, mr1, P<const float>(p[14]), ldd, R, S, d, training, dY, P<float>(p[21]),
                         P<float>(p[22]), stats + (size_t)(2 * l) * stats_stride, st);
    if (rc) return rc;
    // second Linear: dW1, db1 (its input relu(bn0(H)) is recomputed in the prologue), then dH = dY W1
    rc = sb_linear_wgrad(dY, ldd, H, ldh, R, S, d, h, 2, a0, c0, P<float>(p[18]), h, 1, P<float>(p[19]), 0, ws, stream);
    if (rc) return rc;
    rc = sb_linear_fwd(dY, ldd, P<const float>(p[12]), 1, h, nullptr, dH, ldh, R, S, d, h, 0, nullptr, nullptr, 0,
                       nullptr, 0, stream);
    if (rc) return rc;
    // inner BN + ReLU (in place)
    rc = bn_backward(dH, H, a0, c0, mr0, P<const float>(p[11]), ldh, R, S, h, training, dH, P<float>(p[16]),
                     P<float>(p[17]), stats + (size_t)(2 * l + 1) * stats_stride, st);
    if (rc) return rc;
    // first Linear: dW0, dA = dH W0
    rc = sb_linear_wgrad(dH, ldh, A, ld_in, R, S, h, d_in, 0, nullptr, nullptr, P<float>(p[15]), d_in, 1, nullptr, 0,
                         ws, stream);
    if (rc) return rc;
    rc = sb_linear_fwd(dH, ldh, P<const float>(p[10]), 1, d_in, nullptr, dA, ld_in, R, S, h, d_in, 0, nullptr, nullptr,
                       0, nullptr, 0, stream);
    if (rc) return rc;
    // transposed aggregate + residual stream + d eps
    if (l > 0) rc = agg(sl, dA, G, G, X, P<double>(p[20]), P<const float>(p[13]), S, ld_in, true, stream);
    else rc = agg(sl, dA, out0, nullptr, X, P<double>(p[20]), P<const float>(p[13]), S, ld_in, true, stream);
    if (rc) return rc;
  }
  return SB_OK;
}

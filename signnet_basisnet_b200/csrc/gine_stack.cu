// The GINE layer stack of the downstream predictor behind TWO C-ABI calls.
//
// Reference: GNN.forward, the loop at Alchemy/sign_net/model.py:44-57 -
//     e = edge_encoder(edge_attr)                       elements.MLP(nfeat_edge, nhid, 1) or DiscreteEncoder (:21-37)
//     x = conv(x, edge_index, e)                        GINEConv (pyg_gnn_wrapper.py:19-28): nn((1+eps) x_i + sum relu(x_j + e_ij)),
//                                                       nn = MLP(nhid, nhid, 2, with_final_activation=False) (elements.py:39-69)
//     x = relu(norm(x)); x = x + previous_x
// Why: this loop is ~13 kernel entry points per layer forward and ~25 backward, each behind its own autograd Function
// and allocations in Python; the reference's own Alchemy configuration (16 such layers on 128 small graphs) is entirely
// host-bound that way (618 C-ABI calls, 19.5 ms of Python for a step the GPU finishes in a few ms,
// scripts/host_probe_cfg2.py).  Here the caller hands over a host table of device pointers and the launch sequence - the
// SAME entry points in the SAME order as signnet_basisnet_b200/model.py issues them one by one - is enqueued back to
// back.  No allocation, no synchronisation, nothing retained.
#include <cuda_runtime.h>
#include "common.cuh"
#include "../../include/signnet_b200.h"

namespace {
template <typename T>
inline T* P(int64_t v) { return reinterpret_cast<T*>(static_cast<uintptr_t>(v)); }

struct Gr {   // graph arrays + sizes shared by all layers
  const int32_t *in_ptr, *in_src, *in_eid, *out_ptr, *out_dst, *out_eid;
  const int64_t* edge_index;
  const void* edge_attr;     // float [E, ld_ea] (continuous) or int64 [E, F] (discrete)
  int32_t* flags;            // embedding range flag (may be null)
  int64_t N, E, ld_ea;
  int d, ld, nfe, F, V;
};
Gr read_gr(const int64_t* p, const int64_t* n) {
  Gr g;
  g.in_ptr = P<const int32_t>(p[0]); g.in_src = P<const int32_t>(p[1]); g.in_eid = P<const int32_t>(p[2]);
  g.out_ptr = P<const int32_t>(p[3]); g.out_dst = P<const int32_t>(p[4]); g.out_eid = P<const int32_t>(p[5]);
  g.edge_index = P<const int64_t>(p[6]); g.edge_attr = P<const void>(p[7]); g.flags = P<int32_t>(p[8]);
  g.N = n[0]; g.E = n[1]; g.ld_ea = n[2]; g.d = (int)n[3]; g.ld = (int)n[4]; g.nfe = (int)n[5]; g.F = (int)n[6];
  g.V = (int)n[7];
  return g;
}

#define GS_BN_SMALL_ROWS 8192   // sb_bn_act_fwd / sb_bn_act_bwd run as ONE kernel up to this many rows (elementwise.cu)

// batch_norm_act of functional.py: column statistics -> finalize (+ running buffers) -> act(a x + c) (+ res).
// `stats`: this BatchNorm's own ZERO-INITIALISED fp64 [2,C] region of the caller's arena - no memset per BatchNorm.
int bn_act(const float* x, float* out, const float* res, int64_t M, int ld, int C, const float* gamma, const float* beta,
           float* rm, float* rv, double* stats, float* a, float* c, double* mr, int training, float mom, float eps,
           void* st) {
  if (M <= GS_BN_SMALL_ROWS)
    return sb_bn_act_fwd(x, ld, M, 1, C, gamma, beta, rm, rv, mom, eps, training, 1, res, out, stats, a, c, mr, st);
  if (training) {
    const int rc = sb_col_stats(x, ld, M, 1, C, stats, st);
    if (rc) return rc;
  }
  return sb_bn_apply_fwd(x, training ? stats : nullptr, M, 1, C, gamma, beta, rm, rv, mom, eps, training, 1, res, out, ld, M,
                         a, c, mr, st);
}
// BatchNormActFn.backward: dz (may alias gout) <- d/dx of relu(BN(x)); dgamma, dbeta
int bn_act_bwd(const float* gout, const float* x, const float* a, const float* c, const double* mr, const float* gamma,
               int64_t M, int ld, int C, int training, float* dz, float* dgamma, float* dbeta, double* stats, double* coef,
               cudaStream_t st) {
  if (M <= GS_BN_SMALL_ROWS)
    return sb_bn_act_bwd(gout, x, a, c, mr, gamma, ld, M, 1, C, 1, training, dz, dgamma, dbeta, stats, coef, (void*)st);
  const int rc = sb_bn_bwd_reduce(gout, x, a, c, mr, nullptr, ld, M, 1, C, 1, stats, st);
  if (rc) return rc;
  return sb_bn_apply_bwd(gout, x, stats, mr, a, c, gamma, M, training, dz, dgamma, dbeta, ld, M, 1, C, st);
}
}  // namespace

#define GS_FWD_COLS 40
#define GS_BWD_COLS 44
#define GS_MAXF 4

// graph_ptrs[9]  = { in_ptr, in_src, in_eid, out_ptr, out_dst, out_eid, edge_index, edge_attr, embedding flags|0 }
// graph_ints[8]  = { N, E, ld_ea (row stride of edge_attr in elements), d, ld, nfe (continuous edge features; 0 = discrete),
//                    F (discrete feature columns, <= 4), V (rows of an embedding table) }
// layer_ptrs[l]  = { X_in, X_out, A, H, Hn, Y, Ee|0, e,   We|0, ge|0, be|0, rme|0, rve|0,   eps, W0, g0, b0, rm0, rv0, W1, g1,
//                    b1, rm1, rv1,   stats_e|0, stats_0|0, stats_1|0 (fp64 [2,d], zeroed by the caller; 0 in eval mode),
//                    ae, ce, mre, a0, c0, mr0, a1, c1, mr1,   table_0 .. table_3 (discrete) }
extern "C" int sb_gine_stack_fwd(const int64_t* layer_ptrs, int32_t L, const int64_t* graph_ptrs, const int64_t* graph_ints,
                                 int32_t training, float momentum, float bn_eps, void* stream) {
  SB_CHECK_ARG(layer_ptrs && graph_ptrs && graph_ints && L >= 1, "sb_gine_stack_fwd: bad arguments");
  const Gr g = read_gr(graph_ptrs, graph_ints);
  SB_CHECK_ARG(g.F <= GS_MAXF && (g.nfe > 0) != (g.F > 0), "sb_gine_stack_fwd: edge features must be continuous or <= 4 discrete");
  for (int l = 0; l < L; ++l) {
    const int64_t* p = layer_ptrs + (size_t)l * GS_FWD_COLS;
    const float* X = P<const float>(p[0]);
    float *Xn = P<float>(p[1]), *A = P<float>(p[2]), *H = P<float>(p[3]), *Hn = P<float>(p[4]), *Y = P<float>(p[5]);
    float *Ee = P<float>(p[6]), *e = P<float>(p[7]);
    int rc;
    if (g.nfe > 0) {   // edge_encoder = Linear(nfeat_edge -> d, no bias) -> BN -> ReLU
      rc = sb_linear_fwd(static_cast<const float*>(g.edge_attr), g.ld_ea, P<const float>(p[8]), g.nfe, 1, nullptr, Ee, g.ld,
                         g.E, 1, g.nfe, g.d, 0, nullptr, nullptr, 0, nullptr, 0, stream);
      if (rc) return rc;
      rc = bn_act(Ee, e, nullptr, g.E, g.ld, g.d, P<const float>(p[9]), P<const float>(p[10]), P<float>(p[11]),
                  P<float>(p[12]), P<double>(p[24]), P<float>(p[27]), P<float>(p[28]), P<double>(p[29]), training, momentum,
                  bn_eps, stream);
      if (rc) return rc;
    } else {           // edge_encoder = DiscreteEncoder: sum of per-column embeddings
      for (int f = 0; f < g.F; ++f) {
        rc = sb_embedding_fwd(static_cast<const int64_t*>(g.edge_attr) + f, g.ld_ea, P<const float>(p[36 + f]), g.V, g.d,
                              g.E, e, g.ld, f > 0, g.flags, stream);
        if (rc) return rc;
      }
    }
    rc = sb_gine_agg_fwd(X, e, P<const float>(p[13]), g.in_ptr, g.in_src, g.in_eid, g.N, g.ld, A, stream);
    if (rc) return rc;
    rc = sb_linear_fwd(A, g.ld, P<const float>(p[14]), g.d, 1, nullptr, H, g.ld, g.N, 1, g.d, g.d, 0, nullptr, nullptr, 0,
                       nullptr, 0, stream);
    if (rc) return rc;
    rc = bn_act(H, Hn, nullptr, g.N, g.ld, g.d, P<const float>(p[15]), P<const float>(p[16]), P<float>(p[17]),
                P<float>(p[18]), P<double>(p[25]), P<float>(p[30]), P<float>(p[31]), P<double>(p[32]), training, momentum,
                bn_eps, stream);
    if (rc) return rc;
    rc = sb_linear_fwd(Hn, g.ld, P<const float>(p[19]), g.d, 1, nullptr, Y, g.ld, g.N, 1, g.d, g.d, 0, nullptr, nullptr, 0,
                       nullptr, 0, stream);
    if (rc) return rc;
    rc = bn_act(Y, Xn, X, g.N, g.ld, g.d, P<const float>(p[20]), P<const float>(p[21]), P<float>(p[22]), P<float>(p[23]),
                P<double>(p[26]), P<float>(p[33]), P<float>(p[34]), P<double>(p[35]), training, momentum, bn_eps, stream);
    if (rc) return rc;
  }
  return SB_OK;
}

// layer_ptrs[l] = { X_in, A, H, Hn, Y, Ee|0, e,   ae, ce, mre, a0, c0, mr0, a1, c1, mr1,   We|0, ge|0, eps, W0, g0, W1, g1,
//                   dWe|0, dge|0, dbe|0, deps (fp64 scalar, zeroed by the caller), dW0, dg0, db0, dW1, dg1, db1,
//                   idx-table grads dtable_0 .. dtable_3 (discrete), reserved... }
// scratch[11]   = { Ga (in: dL/dX_L), Gb (the residual-stream gradient ping-pongs between the two: after L layers dL/dX_0 is
//                   in Ga if L is even, else in Gb), dY, dH, dA, dx, de, stats arena fp64 [3L][2,ld] ZERO-INITIALISED (a region per BatchNorm), coef fp64 [3,d], wgrad workspace,
//                   embedding-backward workspace|0 }
extern "C" int sb_gine_stack_bwd(const int64_t* layer_ptrs, int32_t L, const int64_t* graph_ptrs, const int64_t* graph_ints,
                                 const int64_t* scratch, int32_t training, void* stream) {
  SB_CHECK_ARG(layer_ptrs && graph_ptrs && graph_ints && scratch && L >= 1, "sb_gine_stack_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const Gr g = read_gr(graph_ptrs, graph_ints);
  float *G = P<float>(scratch[0]), *Gn = P<float>(scratch[1]), *dY = P<float>(scratch[2]), *dH = P<float>(scratch[3]),
        *dA = P<float>(scratch[4]), *dx = P<float>(scratch[5]), *de = P<float>(scratch[6]);
  double *stats = P<double>(scratch[7]), *coef = P<double>(scratch[8]);   // stats: arena [3L][2*ld], zero-initialised
  int bn_i = 0;
  float *ws = P<float>(scratch[9]), *ews = P<float>(scratch[10]);
  for (int l = L - 1; l >= 0; --l) {
    const int64_t* p = layer_ptrs + (size_t)l * GS_BWD_COLS;
    const float *X = P<const float>(p[0]), *A = P<const float>(p[1]), *H = P<const float>(p[2]), *Hn = P<const float>(p[3]),
                *Y = P<const float>(p[4]), *Ee = P<const float>(p[5]), *e = P<const float>(p[6]);
    // outer BN + ReLU (+ residual: its gradient is G itself and is added back below)
    int rc = bn_act_bwd(G, Y, P<const float>(p[13]), P<const float>(p[14]), P<const double>(p[15]), P<const float>(p[22]), g.N,
                        g.ld, g.d, training, dY, P<float>(p[31]), P<float>(p[32]), stats + (size_t)(bn_i++) * 2 * g.ld, coef, st);
    if (rc) return rc;
    // second Linear of the MLP: dHn, dW1   (LinearFn.backward: input gradient first, then the weight gradient)
    rc = sb_linear_fwd(dY, g.ld, P<const float>(p[21]), 1, g.d, nullptr, dH, g.ld, g.N, 1, g.d, g.d, 0, nullptr, nullptr, 0,
                       nullptr, 0, stream);
    if (rc) return rc;
    rc = sb_linear_wgrad(dY, g.ld, Hn, g.ld, g.N, 1, g.d, g.d, 0, nullptr, nullptr, P<float>(p[30]), g.d, 1, nullptr, 0, ws,
                         stream);
    if (rc) return rc;
    // inner BN + ReLU (in place)
    rc = bn_act_bwd(dH, H, P<const float>(p[10]), P<const float>(p[11]), P<const double>(p[12]), P<const float>(p[20]), g.N,
                    g.ld, g.d, training, dH, P<float>(p[28]), P<float>(p[29]), stats + (size_t)(bn_i++) * 2 * g.ld, coef, st);
    if (rc) return rc;
    // first Linear: dA, dW0
    rc = sb_linear_fwd(dH, g.ld, P<const float>(p[19]), 1, g.d, nullptr, dA, g.ld, g.N, 1, g.d, g.d, 0, nullptr, nullptr, 0,
                       nullptr, 0, stream);
    if (rc) return rc;
    rc = sb_linear_wgrad(dH, g.ld, A, g.ld, g.N, 1, g.d, g.d, 0, nullptr, nullptr, P<float>(p[27]), g.d, 1, nullptr, 0, ws,
                         stream);
    if (rc) return rc;
    // GINE aggregate: dx (by source), de (per edge), d eps
    rc = sb_gine_agg_bwd(dA, X, e, P<const float>(p[18]), g.edge_index, g.out_ptr, g.out_dst, g.out_eid, g.N, g.E, g.ld, dx,
                         de, P<double>(p[26]), stream);
    if (rc) return rc;
    // residual stream: dL/dX_l = dx + G   (into the other buffer: the element-wise kernel's operands do not alias)
    rc = sb_affine_act_res(dx, nullptr, nullptr, G, Gn, g.ld, g.N, 1, g.ld, 0, stream);
    if (rc) return rc;
    { float* t = G; G = Gn; Gn = t; }
    // edge encoder
    if (g.nfe > 0) {
      rc = bn_act_bwd(de, Ee, P<const float>(p[7]), P<const float>(p[8]), P<const double>(p[9]), P<const float>(p[17]), g.E,
                      g.ld, g.d, training, de, P<float>(p[24]), P<float>(p[25]), stats + (size_t)(bn_i++) * 2 * g.ld, coef, st);
      if (rc) return rc;
      rc = sb_linear_wgrad(de, g.ld, static_cast<const float*>(g.edge_attr), g.ld_ea, g.E, 1, g.d, g.nfe, 0, nullptr, nullptr,
                           P<float>(p[23]), g.nfe, 1, nullptr, 0, ws, stream);
      if (rc) return rc;
    } else {
      for (int f = 0; f < g.F; ++f) {
        rc = sb_embedding_bwd(static_cast<const int64_t*>(g.edge_attr) + f, g.ld_ea, de, g.ld, g.V, g.d, g.E,
                              P<float>(p[33 + f]), ews, stream);
        if (rc) return rc;
      }
    }
  }
  return SB_OK;
}

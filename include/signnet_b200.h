/* libsignnet_b200 — C ABI of the B200-native SignNet hot path.
 *
 * The reference (cptq/SignNet-BasisNet) has no FFI: its boundary for this path is the nn.Module surface
 * (Alchemy/sign_net/sign_net.py:74-132, GraphPrediction/layers/deepsigns.py:33-86) and, beneath it, calls into
 * third-party graph libraries.  Each entry point below replaces one of those library call sites (cited per function);
 * the Python modules in signnet_basisnet_b200/ bind them with ctypes exactly as INTEGRATION.md shows.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch caching allocator); nothing is allocated, retained
 *     or freed here; `stream` is a cudaStream_t (torch.cuda.current_stream().cuda_stream), all work is enqueued on it;
 *   - return value 0 = ok, non-zero = error, message via sb_last_error() (thread-local);
 *   - index data at the API is int64 (`edge_index`, `batch`) as in PyG/DGL; floating point is fp32;
 *   - "slot rows": activations of phi live in the ragged layout  row(b, j, i) = row_ptr[b] + j*n_b + i
 *     (graph b, eigenvector slot j < k_b, local node i < n_b), as a [S, R, ld] tensor (S = 2 sign passes,
 *     ld = feature dim rounded up to 4 floats, padding columns kept at 0).  Only valid slots exist, so the reference's
 *     boolean-mask bookkeeping (sign_net.py:38-39, masked_layers.py:59-60) has no counterpart.
 *   - `G` ("groups") = leading dimension over which BatchNorm statistics are kept separate (the two sign passes).
 *   - process-wide state: one process drives one GPU.  The only mutable globals are the kernel-selection switch
 *     (sb_set_tensor_cores) with its two diagnostics (sb_last_linear_kernel / sb_last_wgrad_kernel) and the one-time
 *     per-kernel cudaFuncSetAttribute; set the switch before launching work from several host threads.
 */
#ifndef SIGNNET_B200_H
#define SIGNNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_ABI_VERSION 1

/* bits of the `flags` word written by the bookkeeping kernels (device int32, caller zero-initialises) */
#define SB_FLAG_BATCH_UNSORTED 1   /* `batch` is not non-decreasing */
#define SB_FLAG_BATCH_RANGE 2      /* `batch` value outside [0, B) */
#define SB_FLAG_EDGE_RANGE 4       /* edge endpoint outside [0, N) */
#define SB_FLAG_EDGE_CROSS_GRAPH 8 /* an edge joins two different graphs of the batch */

const char* sb_last_error(void);
int sb_abi_version(void);
int sb_device_sm_count(void);

/* ---- bookkeeping (bit-exact integer work) ------------------------------------------------------------------------ */

/* graph_ptr[B+1] (int32 node offsets) from the sorted batch vector.
 * Replaces scatter(ones, batch) + cumsum: Alchemy/sign_net/transform.py:28-30, sign_net.py:100. */
int sb_graph_ptr(const int64_t* batch, int64_t N, int32_t B, int32_t* graph_ptr, int32_t* flags, void* stream);

/* Slot-row layout. k_b = masked ? min(n_b,k) : k.  row_ptr[B+1] = prefix(n_b*k_b); vec_ptr[B+1] = prefix(n_b^2)
 * (offsets into the ragged eigen_vectors, transform.py:14); unit_ptr[B+1] = prefix of aggregate work units for
 * `tile_rows`; summary[6] = {R, n_max, max k_b, sum n_b^2, #units, #graphs with n_b > tile_rows}.
 * Replaces the mask construction sign_net.py:100-102 / deepsigns.py:66-78 and to_dense_EVD's index math. */
int sb_slot_layout(const int32_t* graph_ptr, int32_t B, int32_t k, int32_t masked, int32_t tile_rows,
                   int64_t* row_ptr, int64_t* vec_ptr, int32_t* unit_ptr, int64_t* summary, void* stream);
int sb_agg_units(const int32_t* graph_ptr, int32_t B, int32_t k, int32_t masked, int32_t tile_rows,
                 int32_t* unit_ptr, void* stream);
/* nbr_pack[N]: per node, up to four local neighbour ids (id - first node of its graph) of one CSR, a byte each in
 * CSR order, 0xFF = empty; 0xFE in byte 3 = degree > 4 or id > 253 (the aggregate walks the CSR for that node). */
int sb_pack_neighbours(const int64_t* batch, const int32_t* graph_ptr, const int32_t* nbr_ptr, const int32_t* nbr_idx,
                       int64_t N, uint32_t* nbr_pack, void* stream);
/* unit_desc[U][12] int32: one record per aggregate work unit = (graph, slot chunk) tile of <= tile_rows rows:
 * {row_rel lo, hi, n_b, rows, node0, ceil(2^32/n_b), first edge / edge count of the graph in the CSR by destination,
 * the same for the CSR by source, 0, 0}.  cap_units = records allocated (U <= N if masked else B*k).  The producer warp
 * of the TMA aggregate finds its tile with one coalesced load of this record. */
int sb_agg_unit_desc(const int32_t* graph_ptr, const int64_t* row_ptr, const int32_t* unit_ptr, const int32_t* in_ptr,
                     const int32_t* out_ptr, int32_t B, int32_t k, int32_t masked, int32_t tile_rows,
                     int32_t* unit_desc, int64_t cap_units, void* stream);

/* Stable CSR by destination (in_*) and by source (out_*) of edge_index[2,E]; rows keep edge-id order so neighbour
 * sums accumulate in the order torch's CPU index_add_ uses.  workspace: >= 2*(N+1) + 2*ceil((N+1)/4096) int32.
 * Replaces the per-call gather/scatter indexing inside PyG MessagePassing.propagate / DGL update_all. */
int sb_build_csr(const int64_t* edge_index, int64_t E, int64_t N, const int64_t* batch, int32_t* in_ptr,
                 int32_t* in_src, int32_t* in_eid, int32_t* out_ptr, int32_t* out_dst, int32_t* out_eid,
                 int32_t* workspace, int64_t workspace_ints, int32_t* flags, void* stream);

/* phi input x0[2, R]: +/- eigenvector entries in slot-row order, from the ragged per-graph V (row-major [node, eig])
 * or from a dense-list tensor eigvecs[N, >=k].  Replaces to_dense_list_EVD + unsqueeze/transpose + `-x`
 * (transform.py:52-61, sign_net.py:104,113; GNN3d :31). */
int sb_phi_input_ragged(const float* eigen_vectors, const int64_t* batch, const int32_t* graph_ptr,
                        const int64_t* row_ptr, const int64_t* vec_ptr, int64_t N, int32_t k, int32_t masked,
                        int64_t R, float* x0, void* stream);
int sb_phi_input_dense(const float* eigvecs, int64_t ld, const int64_t* batch, const int32_t* graph_ptr,
                       const int64_t* row_ptr, int64_t N, int32_t k, int32_t masked, int64_t R, float* x0,
                       void* stream);
/* eigenvalue feature per slot row (input of SignNet.eigen_encoder, sign_net.py:107-108) */
int sb_slot_eigval(const float* eigen_values, const int64_t* batch, const int32_t* graph_ptr, const int64_t* row_ptr,
                   int64_t N, int32_t k, int32_t masked, float* out, void* stream);
/* to_dense_list_EVD itself (transform.py:52-61): eigS/eigV [N, nmax] zero padded, mask[N, nmax] (any may be NULL) */
int sb_dense_list_evd(const float* eigen_values, const float* eigen_vectors, const int64_t* batch,
                      const int32_t* graph_ptr, const int64_t* vec_ptr, int64_t N, int32_t nmax, float* eigS,
                      float* eigV, uint8_t* mask, void* stream);

/* slot rows <-> the reference's padded dense view [N, k, C] (return value of GNN3d, sign_net.py:44; phi(x)+phi(-x),
 * sign_net.py:113 / deepsigns.py:73).  rows_to_dense sums over the S sign passes; dense_to_rows writes +v (and -v or a
 * copy) and zero-fills the padding columns. */
int sb_rows_to_dense(const float* rows, int64_t ld, int64_t R, int32_t S, const int64_t* batch,
                     const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N, int32_t k, int32_t masked,
                     int32_t C, float* dense, void* stream);
int sb_dense_to_rows(const float* dense, int64_t R, int32_t S, int32_t negate_second, const int64_t* batch,
                     const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N, int32_t k, int32_t masked,
                     int32_t C, int64_t ld, float* rows, void* stream);

/* ---- K1: GIN neighbourhood aggregate ------------------------------------------------------------------------------
 * out = [res +] (1+eps)*x + sum_{nbr} x   on [S, R, ld] slot rows; (nbr_ptr, nbr_idx) = CSR by destination for the
 * forward, by source for the backward; nbr_pack = the packed neighbour words of the SAME CSR (sb_pack_neighbours);
 * unit_desc or nbr_pack NULL => generic kernel.  Optional dot_out += sum(x * dotx) (= d eps in the backward).
 * Replaces gnn.GINConv(Identity(), train_eps=True) (masked_layers.py:70,75) / dgl GINConv(.,'sum') (gnns.py:90-98). */
int sb_gin_agg(const float* x, float* out, const float* res, const float* dotx, double* dot_out, const float* eps,
               const int32_t* graph_ptr, const int32_t* unit_ptr, const int32_t* unit_desc, const uint32_t* nbr_pack,
               const int64_t* row_ptr, const int32_t* nbr_ptr, const int32_t* nbr_idx, int64_t R, int32_t B, int32_t k,
               int32_t masked, int32_t S, int32_t ld, int32_t tile_rows, int32_t force_generic, void* stream);
int sb_gin_agg_tile_rows(int32_t ld); /* rows per shared-memory tile of the TMA path (0 = generic path only) */

/* ---- K1 + K2 fused (csrc/gin_lin_fused.cu): A = sb_gin_agg(X) and H = A W^T (+ fp64 column statistics of H, layout of
 * sb_linear_fwd's `stats`) in one launch - MaskedGINConv.forward = GINConv aggregate then the first Linear of its
 * MaskedMLP (Alchemy/sign_net/model_utils/masked_layers.py:74-84, :54-58).  The aggregated tile goes from shared memory
 * straight into the tcgen05 contraction (weight resident in tensor memory); A is still written once because the
 * backward's weight gradient reads it.  Fast path: ld = K = 128, h <= 128, S <= 2, TMA tile layout (tile_rows = 64, not
 * `generic`).  Anything else returns SB_ERR_UNSUPPORTED (3) WITHOUT setting an error: run sb_gin_agg + sb_linear_fwd,
 * which produce bit-identical A and H.  sb_set_fused_agg_linear(0/1) (env SB_FUSED_AGG_LINEAR) switches the fast path
 * off/on and returns the previous setting (-1 = undecided). */
int sb_gin_linear_fused_fwd(const float* x, float* a_out, float* h_out, double* stats, const float* eps, const float* w,
                            int64_t w_rs, int64_t w_cs, int32_t K, int32_t h, int64_t ldh, const int32_t* unit_ptr,
                            const int32_t* unit_desc, const uint32_t* nbr_pack, const int32_t* nbr_ptr,
                            const int32_t* nbr_idx, int64_t R, int32_t B, int32_t S, int32_t ld, int32_t tile_rows,
                            int32_t generic, void* stream);
int sb_set_fused_agg_linear(int32_t enable);

/* ---- the whole phi stack behind two calls (csrc/phi_stack.cu) -------------------------------------------------------
 * GNN3d.forward (Alchemy/sign_net/sign_net.py:28-44) for both sign passes: L x { sb_gin_agg -> sb_linear_fwd (+ column
 * statistics) -> sb_bn_finalize -> sb_linear_fwd (BN + ReLU prologue, + statistics) -> sb_bn_finalize ->
 * sb_affine_act_res }, and its backward, enqueued back to back on `stream` by host C++ instead of ~20 Python-driven
 * calls per layer.  The tables are HOST arrays of device pointers / sizes (read during the call, not retained):
 *   fwd layer_ptrs[L][25] = { X_in, A, H, Y, X_out,  W0, bn0.weight, bn0.bias, W1, b1|NULL, eps, bn.weight, bn.bias,
 *                             bn0.running_mean, bn0.running_var, bn.running_mean, bn.running_var,
 *                             st0|NULL, st1|NULL (fp64 [S,2,C], zeroed by the caller; NULL = eval mode),
 *                             a0, c0, mr0, a1, c1, mr1 (outputs of sb_bn_finalize, kept for the backward) }
 *   bwd layer_ptrs[L][23] = { X, A, H, Y, a0, c0, mr0, a1, c1, mr1,  W0, bn0.weight, W1, eps, bn.weight,
 *                             dW0, dbn0.weight, dbn0.bias, dW1, db1|NULL, deps (fp64 scalar, zeroed by the caller),
 *                             dbn.weight, dbn.bias }
 *   dims[L][4]            = { d_in, h, d, ld_in }  (ld_in = 1 for the [S,R] input of the first layer)
 *   slot_ptrs[2][10]      = { graph_ptr, unit_ptr, unit_desc, in_pack, out_pack, row_ptr, in_ptr, in_src, out_ptr,
 *                             out_dst }; slot_ints[2][6] = { R, B, k, masked, tile_rows, generic }: row 0 = the layout
 *                             for padded rows, row 1 = the layout used while ld_in % 4 != 0 (first layer)
 *   scratch[8]            = { G (in: dL/dX_L, updated in place), dY, dH, dA, out0, stats fp64 [S,2,Cmax],
 *                             coef fp64 [3,S,Cmax], sb_linear_wgrad workspace } */
int sb_phi_stack_fwd(const int64_t* layer_ptrs, const int32_t* dims, int32_t L, const int64_t* slot_ptrs,
                     const int64_t* slot_ints, int32_t S, int32_t training, float momentum, float bn_eps, void* stream);
int sb_phi_stack_bwd(const int64_t* layer_ptrs, const int32_t* dims, int32_t L, const int64_t* slot_ptrs,
                     const int64_t* slot_ints, const int64_t* scratch, int32_t S, int32_t training, void* stream);

/* ---- K2: Linear with fused BatchNorm prologue / statistics epilogue ------------------------------------------------
 * y[g*R+r, n] (+)= sum_k f(x[g*R+r, k]) * W[n*w_rs + k*w_cs] + bias[n];  f: pro 0 none, 1 pa*x+pc, 2 relu(pa*x+pc)
 * with pa/pc [G, K]; optional relu on the output; stats[G,2,N] += column sum / sum of squares of the output (fp64).
 * Columns N..ldy-1 of y are zero-filled.  Replaces nn.Linear + mask writes in MaskedMLP (masked_layers.py:54-64),
 * layers/mlp.py:37-56, elements.py:57-65. */
int sb_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t w_rs, int64_t w_cs, const float* bias,
                  float* y, int64_t ldy, int64_t R, int32_t G, int32_t K, int32_t N, int32_t pro, const float* pa,
                  const float* pc, int32_t relu, double* stats, int32_t accumulate, void* stream);
/* 1 (default): contractions with 16 <= K,N <= 128 run on tcgen05 (3xTF32; linear_tc.cu, wgrad_tc_tma.cu for N = K = 128,
 * wgrad_tc.cu otherwise); 2: the same with the register-fed wgrad_tc.cu for every weight gradient (A/B switch of the
 * tests); 0: fp32 FFMA everywhere.  Returns the previous setting (-1 = not yet decided; env SB_DISABLE_TC=1 = 0). */
int sb_set_tensor_cores(int32_t enable);
/* 1 (default): problems with <= 8 192 rows (all groups) run on the small-row FFMA kernel (32-row tiles, both operands
 * streamed; the predictor / rho contractions); 0: they take the tcgen05 / 128-row kernels like everything else (A/B
 * switch of the tests).  Returns the previous setting. */
int sb_set_small_rows(int32_t enable);
/* Which kernel the last block launch of sb_linear_fwd used: 0 FFMA (128-row tiles), 1 tcgen05, 4 FFMA small-row,
 * -1 none yet (the rank-1 / row-dot streaming kernels do not update it).  Diagnostics for the tests. */
int sb_last_linear_kernel(void);
/* Same for the last block launch of sb_linear_wgrad: 0 FFMA, 1 tcgen05 register-fed (wgrad_tc.cu), 3 tcgen05 TMA-fed
 * (wgrad_tc_tma.cu). */
int sb_last_wgrad_kernel(void);
/* dw[n*rs + k*cs] (+)= sum gy[., n] * f(x[., k]);  db[n] (+)= sum gy[., n]   (deterministic two-stage reduction) */
int sb_linear_wgrad(const float* gy, int64_t ldg, const float* x, int64_t ldx, int64_t R, int32_t G, int32_t N,
                    int32_t K, int32_t pro, const float* pa, const float* pc, float* dw, int64_t dw_rs,
                    int64_t dw_cs, float* db, int32_t accumulate, float* workspace, void* stream);
int64_t sb_linear_wgrad_workspace_floats(void);

/* ---- K3/K4: BatchNorm over slot rows + element-wise glue -----------------------------------------------------------
 * Replaces MaskedBN (masked_layers.py:13-20) / nn.BatchNorm1d (gnns.py:107-112, mlp.py:42-46). */
int sb_col_stats(const float* x, int64_t ld, int64_t R, int32_t G, int32_t C, double* stats, void* stream);
int sb_bn_finalize(const double* stats, int64_t M, int32_t G, int32_t C, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float momentum, float eps, int32_t training, float* a,
                   float* c, double* mean_rstd /*[2,G,C] fp64 mean and 1/sqrt(var+eps), kept for the backward*/,
                   void* stream);
/* out = act(pa*y + pc) + res  (GNN3d tail: norm -> relu -> residual, sign_net.py:40-43) */
int sb_affine_act_res(const float* y, const float* pa, const float* pc, const float* res, float* out, int64_t ld,
                      int64_t R, int32_t G, int32_t C, int32_t relu, void* stream);
/* backward of act(BN(y)): dz = gout*[pa*y+pc > 0] (written only if dz != NULL); stats[G,2,C] += (sum dz, sum dz*y_hat)
 * in fp64 */
int sb_bn_bwd_reduce(const float* gout, const float* y, const float* pa, const float* pc, const double* mean_rstd,
                     float* dz, int64_t ld, int64_t R, int32_t G, int32_t C, int32_t relu, double* stats,
                     void* stream);
/* dgamma/dbeta and the fp64 coefficients coef[3,G,C] of  dY = al*dZ + be*(Y - mean) + ga */
int sb_bn_bwd_finalize(const double* stats, int64_t M, int32_t G, int32_t C, const float* gamma,
                       const double* mean_rstd, int32_t training, int32_t accumulate, float* dgamma, float* dbeta,
                       double* coef, void* stream);
/* out = al*dz + be*(t2 - mean) + ga with dz = t1, or dz = t1 * [pa*t2 + pc > 0] when pa/pc are given (the ReLU mask
 * of the forward recomputed here, so sb_bn_bwd_reduce need not write dz: one activation-sized write less per BN). */
int sb_affine2(const float* t1, const float* t2, const double* coef, const double* mean_rstd, const float* pa,
               const float* pc, float* out, int64_t ld, int64_t R, int32_t G, int32_t C, void* stream);

int sb_relu_bwd(const float* g, const float* y, float* out, int64_t n, void* stream); /* out = g * [y > 0] */

/* sb_bn_finalize + sb_affine_act_res in ONE launch (every CTA derives the coefficients from stats[G,2,C] itself, CTA 0
 * publishes a, c, mean_rstd and moves the running buffers; identical values), and sb_bn_bwd_finalize + sb_affine2 in one
 * launch (coefficients from the two backward sums inside the apply kernel; writes dgamma, dbeta).  M = rows per group the
 * statistics were taken over.  Two launches fewer per BatchNorm and step: the launch queue holds ~1 000 launches, i.e.
 * how far the issuing thread may fall behind before the GPU idles. */
int sb_bn_apply_fwd(const float* y, const double* stats, int64_t M, int32_t G, int32_t C, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                    int32_t training, int32_t relu, const float* res, float* out, int64_t ld, int64_t R, float* a, float* c,
                    double* mean_rstd, void* stream);
int sb_bn_apply_bwd(const float* gout, const float* y, const double* stats, const double* mean_rstd, const float* pa,
                    const float* pc, const float* gamma, int64_t M, int32_t training, float* dz, float* dgamma,
                    float* dbeta, int64_t ld, int64_t R, int32_t G, int32_t C, void* stream);

/* BatchNorm + activation (+ residual) as ONE call: column statistics (training) -> a, c, mean_rstd (+ running buffers) ->
 * out = act(a*x + c) (+ res).  nn.BatchNorm1d + ReLU + residual of elements.py:57-65, model.py:41-47,
 * transformer_module.py / sign_net.py:70 on [M, ld] tensors.  For G = 1 and M <= 8 192 rows this is ONE kernel (a CTA owns
 * four channels end to end; same arithmetic as the three streaming kernels, bit-identical results); otherwise it enqueues
 * sb_col_stats -> sb_bn_finalize -> sb_affine_act_res.  stats: fp64 [G,2,C] scratch (zeroed here).
 * sb_set_small_bn(0/1): A/B switch of the tests, returns the previous setting. */
int sb_bn_act_fwd(const float* x, int64_t ld, int64_t M, int32_t G, int32_t C, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, int32_t training, int32_t relu,
                  const float* res, float* out, double* stats, float* a, float* c, double* mean_rstd, void* stream);
/* its backward: dz (may alias gout) = d/dx, dgamma, dbeta.  stats fp64 [G,2,C], coef fp64 [3,G,C]: scratch of the
 * streaming path (sb_bn_bwd_reduce -> sb_bn_bwd_finalize -> sb_affine2). */
int sb_bn_act_bwd(const float* gout, const float* x, const float* a, const float* c, const double* mean_rstd,
                  const float* gamma, int64_t ld, int64_t M, int32_t G, int32_t C, int32_t relu, int32_t training,
                  float* dz, float* dgamma, float* dbeta, double* stats, double* coef, void* stream);
int sb_set_small_bn(int32_t enable);

/* sum over eigenvector slots and sign passes -> [N, ldo]  (sign_net.py:113 + :70 ; deepsigns.py:72-81) */
int sb_slot_sum_fwd(const float* x, int64_t ld, int64_t R, int32_t S, const int64_t* batch, const int32_t* graph_ptr,
                    const int64_t* row_ptr, int64_t N, int32_t k, int32_t masked, int32_t limit_by_n, float* out,
                    int64_t ldo, int32_t C, void* stream);
int sb_slot_sum_bwd(const float* gout, int64_t ldo, float* gx, int64_t ld, int64_t R, int32_t S,
                    const int64_t* batch, const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N, int32_t k,
                    int32_t masked, int32_t limit_by_n, int32_t C, void* stream);

/* ---- K5/K6: GINE predictor pieces on [N, ld] node rows ---------------------------------------------------------------
 * out_i = (1+eps) x_i + sum_{(j->i)} relu(x_j + e_ji), e [E, ld] indexed by edge id.  Replaces gnn.GINEConv
 * (Alchemy/sign_net/model_utils/pyg_gnn_wrapper.py:19-28).  Backward: dx (by source, CSC), de per edge, deps. */
int sb_gine_agg_fwd(const float* x, const float* e, const float* eps, const int32_t* in_ptr, const int32_t* in_src,
                    const int32_t* in_eid, int64_t N, int32_t ld, float* out, void* stream);
int sb_gine_agg_bwd(const float* dA, const float* x, const float* e, const float* eps, const int64_t* edge_index,
                    const int32_t* out_ptr, const int32_t* out_dst, const int32_t* out_eid, int64_t N, int64_t E,
                    int32_t ld, float* dx, float* de, double* deps, void* stream);
/* ---- the GINE layer stack of the predictor behind two calls (csrc/gine_stack.cu) ---------------------------------------
 * The loop of GNN.forward (Alchemy/sign_net/model.py:44-57): L x { edge_encoder(edge_attr) -> GINEConv (aggregate + 2-layer
 * MLP with BN/ReLU in between) -> BN -> ReLU -> + previous }, and its backward: the same entry points in the same order as
 * the Python modules issue them one by one, enqueued back to back by host C++.  HOST tables (read during the call):
 *   graph_ptrs[9] = { in_ptr, in_src, in_eid, out_ptr, out_dst, out_eid, edge_index, edge_attr, embedding flags|NULL }
 *   graph_ints[8] = { N, E, ld_ea (row stride of edge_attr in elements), d, ld, nfe (continuous edge features; 0 = discrete),
 *                     F (discrete edge-feature columns, <= 4), V (rows of an embedding table) }
 *   fwd layer_ptrs[L][40] = { X_in, X_out, A, H, Hn, Y, Ee|0, e,  We|0, bn_e.weight|0, bn_e.bias|0, bn_e.running_mean|0,
 *                     bn_e.running_var|0,  eps, W0, bn0.weight, bn0.bias, bn0.running_mean, bn0.running_var, W1, bn.weight,
 *                     bn.bias, bn.running_mean, bn.running_var,  stats_e|0, stats_0|0, stats_1|0 (fp64 [2,d], zeroed by the
 *                     caller; 0 in eval mode),  ae, ce, mre, a0, c0, mr0, a1, c1, mr1 (sb_bn_finalize outputs),
 *                     table_0 .. table_3 (discrete edge encoder) }
 *   bwd layer_ptrs[L][44] = { X_in, A, H, Hn, Y, Ee|0, e,  ae, ce, mre, a0, c0, mr0, a1, c1, mr1,  We|0, bn_e.weight|0, eps, W0,
 *                     bn0.weight, W1, bn.weight,  dWe|0, dbn_e.weight|0, dbn_e.bias|0, deps (fp64 scalar, zeroed by the caller),
 *                     dW0, dbn0.weight, dbn0.bias, dW1, dbn.weight, dbn.bias,  dtable_0 .. dtable_3, (unused) }
 *   scratch[11]   = { Ga (in: dL/dX_L), Gb (ping-pong: dL/dX_0 ends in Ga if L is even, else Gb), dY, dH, dA, dx, de,
 *                     stats fp64 [2,d], coef fp64 [3,d], sb_linear_wgrad workspace, sb_embedding_bwd workspace|0 } */
int sb_gine_stack_fwd(const int64_t* layer_ptrs, int32_t L, const int64_t* graph_ptrs, const int64_t* graph_ints,
                      int32_t training, float momentum, float bn_eps, void* stream);
int sb_gine_stack_bwd(const int64_t* layer_ptrs, int32_t L, const int64_t* graph_ptrs, const int64_t* graph_ints,
                      const int64_t* scratch, int32_t training, void* stream);

/* ---- K10: edge-gated aggregate of the GatedGCN predictor (GraphPrediction/layers/gatedgcn_layer.py:48-54: dgl
 * apply_edges(u_add_v) + update_all(u_mul_e, sum) + update_all(copy_e, sum)) on [N, ld] node rows / [E, ld] edge rows:
 * e_out_k = (Dh[src_k] + Eh[dst_k]) + Ce_k;  h_out_i = Ah_i + sum_in(Bh[src] * sigmoid(e_out)) / (sum_in sigmoid(e_out)
 * + 1e-6); ss_out / ssh_out [N, ld] keep the two sums for the backward.  Backward: given dh [N, ld] and de [E, ld] (may
 * be NULL) writes dBh, dDh, dEh [N, ld] and dCe [E, ld] (= d e_out); dAh = dh.  Deterministic (CSR / CSC order). */
int sb_gated_agg_fwd(const float* Ah, const float* Bh, const float* Dh, const float* Eh, const float* Ce,
                     const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid, int64_t N, int32_t ld,
                     float* e_out, float* h_out, float* ss_out, float* ssh_out, void* stream);
int sb_gated_agg_bwd(const float* dh, const float* de, const float* Bh, const float* e_new, const float* ss,
                     const float* ssh, const int64_t* edge_index, const int32_t* in_ptr, const int32_t* in_eid,
                     const int32_t* out_ptr, const int32_t* out_dst, const int32_t* out_eid, int64_t N, int64_t E,
                     int32_t ld, float* dBh, float* dDh, float* dEh, float* dCe, void* stream);
/* ---- K12: multi-aggregator reduction of the PNA predictor (GraphPrediction/layers/pna_layer.py:37-68 pretrans_edges +
 * update_all(reduce_func_for_h); pna_utils.py:12-31 aggregators, :73-84 scalers).  Message of edge k: j -> i is
 * (U[j] + V[i]) + Q[k] (the pre-transformation Linear split by input block).  Z [N, ldz >= 13 C], tower-major:
 * Z[i, t*13*tin + (0..tin)] = h[i, t*tin + .], then for scaler s in (identity, amplification, attenuation) and
 * aggregator a in (mean, max, min, std): Z[i, t*13*tin + tin + (s*4 + a)*tin + j].  avg_log = net_params['avg_d']['log'].
 * Backward: dU, dV [N, ld], dQ [E, ld], dh [N, ldh] from dZ (max / min route to the first edge attaining them). */
int sb_pna_agg_fwd(const float* U, const float* V, const float* Q, const float* h, const int32_t* in_ptr,
                   const int32_t* in_src, const int32_t* in_eid, int64_t N, int32_t C, int32_t tin, int64_t ld,
                   int64_t ldh, int64_t ldz, float avg_log, float* Z, void* stream);
int sb_pna_agg_bwd(const float* dZ, const float* U, const float* V, const float* Q, const int32_t* in_ptr,
                   const int32_t* in_src, const int32_t* in_eid, const int32_t* out_ptr, const int32_t* out_eid,
                   int64_t N, int32_t C, int32_t tin, int64_t ld, int64_t ldh, int64_t ldz, float avg_log, float* dU,
                   float* dV, float* dQ, float* dh, void* stream);
/* element-wise companions of the PNA layer: graph normalisation out[r, :] = x[r, :] * s[r] (pna_layer.py:73-74; its own
 * backward) and LeakyReLU (mixing network, pna_utils.py FCLayer): g == NULL -> out = leaky_relu(x), else out = g * act'(x) */
int sb_row_scale(const float* x, const float* s, int64_t M, int64_t ld, float* out, void* stream);
int sb_leaky_relu(const float* g, const float* x, int64_t n, float slope, float* out, void* stream);
/* ---- K13: edge-modulated sparse attention of the graph-Transformer predictor (GraphPrediction/layers/transformer.py:
 * 160-192, full_graph=False: apply_edges(src_dot_dst, scaling, imp_exp_attn, exp) + 2 x send_and_recv(sum)):
 * a_k = sum_c ((K[src,c] Q[dst,c]) / sqrt(d)) E[k,c] per head, s = exp(clamp(a, -5, 5)),
 * out[i] = sum_in(s V[src]) / (sum_in s + 1e-6).  H heads of width d <= 32, H*d <= ld.  araw [E, H] and z [N, H] are kept
 * for the backward; dKe / dVe [E, ld] are scratch (per-edge contributions summed per source node in CSC order). */
int sb_edge_attention_fwd(const float* Q, const float* K, const float* Ef, const float* V, const int32_t* in_ptr,
                          const int32_t* in_src, const int32_t* in_eid, int64_t N, int32_t H, int32_t d, int64_t ld,
                          float* out, float* araw, float* z, void* stream);
int sb_edge_attention_bwd(const float* dout, const float* out, const float* Q, const float* K, const float* Ef,
                          const float* V, const float* araw, const float* z, const int32_t* in_ptr, const int32_t* in_src,
                          const int32_t* in_eid, const int32_t* out_ptr, const int32_t* out_eid, int64_t N, int32_t H,
                          int32_t d, int64_t ld, float* dQ, float* dK, float* dE, float* dV, float* dKe, float* dVe,
                          void* stream);
/* K11: the `canonical` sign convention of train/train_ZINC_graph_regression.py:26-42 (PE baseline): per graph and
 * column flip the sign when the column has fewer non-negative than negative entries or less non-negative mass. */
int sb_canonical_sign(const float* pe, int64_t ldp, const int32_t* graph_ptr, int64_t B, int32_t k, float* out,
                      int64_t ldo, void* stream);
/* graph read-out: scatter(x, batch, reduce='add'|'mean') (model.py:58-61) on the sorted batch */
int sb_segment_pool_fwd(const float* x, int64_t ldx, const int32_t* graph_ptr, int32_t B, int32_t C, int32_t mean,
                        float* out, int64_t ldo, void* stream);
int sb_segment_pool_bwd(const float* gout, int64_t ldo, const int64_t* batch, const int32_t* graph_ptr, int64_t N,
                        int32_t C, int32_t mean, float* gx, int64_t ldx, void* stream);
/* DiscreteEncoder (elements.py:21-37), one integer feature column per call.  An index outside [0, V) sets flags bit 0
 * (if flags != NULL), prints the offending row and TRAPS (the CUDA counterpart of nn.Embedding's IndexError: a
 * device-side assert; the next CUDA call of the process fails). */
int sb_embedding_fwd(const int64_t* idx, int64_t stride, const float* table, int32_t V, int32_t C, int64_t M,
                     float* out, int64_t ldo, int32_t accumulate, int32_t* flags, void* stream);
int sb_embedding_bwd(const int64_t* idx, int64_t stride, const float* g, int64_t ldg, int32_t V, int32_t C, int64_t M,
                     float* dtable, float* workspace, void* stream);
int64_t sb_embedding_bwd_workspace_floats(int32_t V, int32_t C);

/* ---- K7: rho = SetTransformer pieces (Alchemy/sign_net/model_utils/transformer_module.py:44-102) -------------------
 * Self-attention of every node over its k_b valid eigenvector slots (tokens = slot rows of the node), n_head heads of
 * width dk packed in columns [h*dk, (h+1)*dk) of q/k/v/o [R, ld].  scores = (q / temperature) k^T, softmax over the
 * valid keys (== the reference's -1e10 fill + mask), optional attention dropout (counter-based, regenerated in the
 * backward from `seed`), o = P v. */
int sb_attention_fwd(const float* q, const float* k, const float* v, int64_t ld, const int64_t* batch,
                     const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N, int32_t kslots, int32_t masked,
                     int32_t kmax, int32_t n_head, int32_t dk, float temperature, float drop_p, int64_t seed, float* o,
                     void* stream);
int sb_attention_bwd(const float* q, const float* k, const float* v, const float* go, int64_t ld,
                     const int64_t* batch, const int32_t* graph_ptr, const int64_t* row_ptr, int64_t N,
                     int32_t kslots, int32_t masked, int32_t kmax, int32_t n_head, int32_t dk, float temperature,
                     float drop_p, int64_t seed, float* gq, float* gk, float* gv, void* stream);
/* 1 (default): d_k = 32, k_b <= 40, 16-byte aligned rows run on the tensor-core kernels of csrc/attention_mma.cu
 * (warp per (node, head), mma.sync m16n8k8 3xTF32, fragments straight from global memory); 0: the FFMA kernels of
 * csrc/attention_fast.cu (which every other shape takes anyway).  Returns the previous setting (A/B switch of the tests). */
int sb_set_attention_mma(int32_t enable);
/* y = LayerNorm(a + b) * w + beta per row (MaskedLN, masked_layers.py:22-32, eps 1e-6); xsum = a + b and
 * stat[R,2] = (mean, rstd) are kept for the backward; dwb[2,C] (fp64) accumulates (dw, dbeta). */
int sb_layernorm_fwd(const float* a, const float* b, const float* w, const float* beta, int64_t ld, int64_t R,
                     int32_t C, float eps, float* y, float* xsum, float* stat, void* stream);
int sb_layernorm_bwd(const float* g, const float* x, const float* stat, const float* w, int64_t ld, int64_t R,
                     int32_t C, float* dx, double* dwb, void* stream);

/* ---- K8: batched Laplacian eigendecomposition (the step before the path) -------------------
 * Per graph: L = I - D^-1/2 A D^-1/2 (A symmetrised, de-duplicated, no self loops; isolated nodes D^-1/2 := 0) from the
 * CSR by destination, diagonalised by a warp-per-graph parallel cyclic Jacobi (n_b <= 64).  eigen_values[N] ascending
 * per graph, eigen_vectors[sum n_b^2] row-major V[node, eig] at vec_ptr[b] - the layout EVDTransform('sym') produces
 * with torch.linalg.eigh on the CPU (Alchemy/sign_net/transform.py:7-23).  flags bit 0: a graph exceeded nmax. */
int sb_laplacian_evd(const int32_t* graph_ptr, const int32_t* in_ptr, const int32_t* in_src, const int64_t* vec_ptr,
                     int32_t B, int32_t nmax, float* eigen_values, float* eigen_vectors, int32_t* flags, void* stream);

/* ---- K9: BasisNet IGN phi (LearningFilters/ign.py:344-374 contractions_2_to_1, normalization 'inf') ------------------
 * ops[(e*n + i)*ldo + 0..4] = { P_ii, tr(P)/n, sum_j P_ij/n, sum_j P_ji/n, sum_ij P_ij/n^2 } of the eigenspace projectors
 * P_e = V_e V_e^T, columns 5..ldo-1 zero.  _factors reads only the eigenvector blocks V[:, col0[e] .. col0[e]+mult)
 * (4 n mult bytes per eigenspace; the projector the reference materialises in training.py:59-61 never exists);
 * _projectors takes the reference's materialised input P [b, n, n] (workspace: 2 b doubles). */
int sb_ign2to1_ops_factors(const float* V, int64_t ldv, int32_t n, const int32_t* col0, int32_t b, int32_t mult,
                           float* ops, int32_t ldo, void* stream);
int sb_ign2to1_ops_projectors(const float* P, int32_t n, int32_t b, float* ops, int32_t ldo, double* workspace,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIGNNET_B200_H */

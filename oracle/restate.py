"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A functional, plain-torch (CPU, fp32) restatement of the reference's SignNet hot path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this file; the
product package `signnet_basisnet_b200` never does (its ops raise if the CUDA library is missing).

Pinned: the reference ships no test or golden vector for this path, so the pin is against the
reference's OWN modules run in the authoring container: `tests/test_oracle_vs_reference.py` runs every function here
against the unmodified classes imported from /root/reference (oracle/ref_loader.py) on seeded inputs, and
`oracle/make_golden.py` stores reference outputs/gradients as fixtures under tests/golden/ which both this oracle
(`-m "not gpu"`) and the CUDA path (`-m gpu`) are checked against on the GPU box where /root/reference is absent.

Every function takes the model's `state_dict` (reference key names) plus a key prefix, so the same
tensors drive the reference module, this oracle and the CUDA modules.  Integer bookkeeping is numpy / int64 torch and
must match bit-for-bit; floating point is fp32 with tolerance 1e-5 relative (BASELINE.json north_star).

Third-party arithmetic restated here because its source is not under /root/reference:
  torch_geometric 2.0.1 GINConv / GINEConv (pin: Alchemy/setup.sh:3-5), torch_scatter.scatter, dgl GINConv (unpinned).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# Test hook: when set to a callable f(z, tag) -> activation, every ReLU of the oracle goes through it.  The GPU parity
# tests use it to impose the activation pattern observed on the CUDA path (x * mask instead of relu(x)): ReLU is
# discontinuous in its derivative, so a pre-activation that lies within rounding distance of 0 (|z| ~ 1e-7 of scale)
# may land on different sides in two correct fp32 implementations, and a single such flip moves per-channel gradient
# sums of a small test problem by ~1e-3.  With the pattern pinned, gradients are comparable at the 1e-5 bar.
ACTIVATION_OVERRIDE = None


def _relu(z, tag=None):
    if ACTIVATION_OVERRIDE is not None:
        return ACTIVATION_OVERRIDE(z, tag)
    return F.relu(z)


# --------------------------------------------------------------------------------------------------------------------
# integer / layout bookkeeping (bit-exact rows a1, a2)
# --------------------------------------------------------------------------------------------------------------------
def graph_sizes(batch: torch.Tensor, num_graphs: int | None = None) -> np.ndarray:
    """n_b per graph.  Follows `scatter(ones, batch)` sign_net.py:100 / transform.py:29."""
    b = batch.numpy()
    B = int(b.max()) + 1 if num_graphs is None else num_graphs
    return np.bincount(b, minlength=B).astype(np.int64)


def dense_list_evd(eig_vals: torch.Tensor, eig_vecs: torch.Tensor, batch: torch.Tensor):
    """Ragged per-graph EVD -> dense-list layout.  Follows to_dense_EVD + to_dense_list_EVD
    (Alchemy/sign_net/transform.py:26-49, :52-61) without the [B,Nmax,Nmax] detour:
        eigS[i, j] = eigenvalue j of graph(i)         (j < n_b, else 0)
        eigV[i, j] = V_b[local(i), j]                 (row-major flattened V, transform.py:14)
    """
    n = graph_sizes(batch)
    nmax = int(n.max())
    N = batch.numel()
    node_ptr = np.concatenate([[0], np.cumsum(n)])
    vec_ptr = np.concatenate([[0], np.cumsum(n * n)])
    ev, evec = eig_vals.numpy(), eig_vecs.numpy()
    S = np.zeros((N, nmax), dtype=ev.dtype)
    V = np.zeros((N, nmax), dtype=evec.dtype)
    for b in range(len(n)):
        nb, p = int(n[b]), int(node_ptr[b])
        S[p:p + nb, :nb] = ev[p:p + nb][None, :]
        V[p:p + nb, :nb] = evec[vec_ptr[b]:vec_ptr[b] + nb * nb].reshape(nb, nb)
    return torch.from_numpy(S), torch.from_numpy(V)


def slot_mask(batch: torch.Tensor, k: int) -> torch.Tensor:
    """mask_full[i, j] = j < n_{batch[i]}  (sign_net.py:100-102; DGL twin deepsigns.py:66-78)."""
    n = torch.from_numpy(graph_sizes(batch))
    return torch.arange(k)[None, :] < n[batch][:, None]


# --------------------------------------------------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------------------------------------------------
def _bn(x2d, sd, p, training):
    """nn.BatchNorm1d on [M, C] rows: batch stats (biased var) + running update (unbiased, momentum .1) in training,
    running stats in eval.  Buffers in `sd` are updated in place like the module would."""
    rm, rv = sd.get(p + "running_mean"), sd.get(p + "running_var")
    use_batch = training or rm is None
    y = F.batch_norm(x2d, None if rm is None else rm, None if rv is None else rv, sd[p + "weight"], sd[p + "bias"],
                     use_batch, BN_MOMENTUM, BN_EPS)
    if training and (p + "num_batches_tracked") in sd:
        sd[p + "num_batches_tracked"] += 1
    return y


def gin_aggregate(x, edge_index, eps):
    """K1.  PyG 2.0.1 GINConv core on node dim -2:  out = scatter_add(x[src] -> dst) ; out += (1+eps)*x
    (masked_layers.py:70,75).  Order of the fp32 adds: edges in edge-id order from zero, then the self term."""
    out = torch.zeros_like(x).index_add_(-2, edge_index[1], x.index_select(-2, edge_index[0]))
    return out + (1 + eps) * x


def _masked_bn(x, mask, sd, p, training):
    """MaskedBN (masked_layers.py:13-20): BN over the valid rows only, other rows untouched."""
    if mask is None:
        sh = x.shape
        return _bn(x.reshape(-1, sh[-1]), sd, p + "bn.", training).reshape(sh)
    out = x.clone()
    out[mask] = _bn(x[mask], sd, p + "bn.", training)
    return out


def masked_mlp(x, mask, sd, p, nlayer=2, with_final_activation=True, training=True, tag=""):
    """MaskedMLP.forward (masked_layers.py:54-64).  The last norm exists in the state_dict but is skipped when
    with_final_activation=False (:61)."""
    for i in range(nlayer):
        x = F.linear(x, sd[f"{p}layers.{i}.weight"], sd.get(f"{p}layers.{i}.bias"))
        if mask is not None:
            x = x * mask.unsqueeze(-1)
        if i < nlayer - 1 or with_final_activation:
            x = _relu(_masked_bn(x, mask, sd, f"{p}norms.{i}.", training), f"{tag}{p}norms.{i}")
    return x


# --------------------------------------------------------------------------------------------------------------------
# PyG flavour: phi = GNN3d of MaskedGINConv (rows a3-a7)
# --------------------------------------------------------------------------------------------------------------------
def gnn3d(x, edge_index, mask, sd, p, n_layer, training=True, tag=""):
    """GNN3d.forward (Alchemy/sign_net/sign_net.py:28-44) on x [N,k,d_in], mask [N,k] -> [N,k,d].
    Per layer: MaskedGINConv (aggregate -> MaskedMLP(2 layers, no final act)) -> zero masked ->
    MaskedBN -> ReLU -> + previous."""
    x = x.transpose(0, 1)
    m = None if mask is None else mask.transpose(0, 1)
    prev = 0
    for l in range(n_layer):
        a = gin_aggregate(x, edge_index, sd[f"{p}convs.{l}.layer.eps"])
        y = masked_mlp(a, m, sd, f"{p}convs.{l}.nn.", 2, False, training, tag)
        if m is not None:
            y = y * m.unsqueeze(-1)
        y = _relu(_masked_bn(y, m, sd, f"{p}norms.{l}.", training), f"{tag}{p}norms.{l}")
        x = y + prev
        prev = x
    return x.transpose(0, 1)


def phi_pm(eigV, edge_index, mask, sd, p, n_layer, training=True):
    """phi(+v) + phi(-v) (sign_net.py:113): two separate passes => per-sign BN batch statistics and two running-stat
    updates per step, +v first."""
    x = eigV.unsqueeze(-1)
    return (gnn3d(x, edge_index, mask, sd, p, n_layer, training, "+") +
            gnn3d(-x, edge_index, mask, sd, p, n_layer, training, "-"))


def _masked_ln(x, mask, sd, p):
    out = x.clone()
    out[mask] = F.layer_norm(x[mask], (x.shape[-1],), sd[p + "ln.weight"], sd[p + "ln.bias"], 1e-6)
    return out


def set_transformer(x, pos, mask, sd, p, n_layer, n_head=4, training=True, attn_dropout=0.0):
    """rho of the PyG trees: SetTransformer.forward (sign_net.py:60-72) with TransformerEncoderLayer
    (transformer_module.py:34-42), MultiHeadAttention (:76-102, post-LN eps 1e-6), masked softmax with -1e10 fill
    (:51-58) and FFN (:113-127).  x [N,k,d], pos [N,k,d] or 0, mask [N,k].
    Reference quirk: MultiHeadAttention builds ScaledDotProductAttention with its default attn_dropout=0.1
    (transformer_module.py:46,85), so in training mode the reference rho is stochastic; parity is checked with that
    dropout at 0 (training) and in eval mode.  `attn_dropout` restates it for the timed CPU baseline."""
    x = x + pos
    N, k, d = x.shape
    dk = d // n_head
    mk = mask.unsqueeze(-1)
    pair = (mask.unsqueeze(1) * mask.unsqueeze(2)).unsqueeze(1)  # [N,1,k,k]
    for l in range(n_layer):
        q = f"{p}transformer_layers.{l}."
        res = x
        Q = F.linear(x, sd[q + "slf_attn.w_qs.weight"]).view(N, k, n_head, dk).transpose(1, 2)
        K = F.linear(x, sd[q + "slf_attn.w_ks.weight"]).view(N, k, n_head, dk).transpose(1, 2)
        V = F.linear(x, sd[q + "slf_attn.w_vs.weight"]).view(N, k, n_head, dk).transpose(1, 2)
        att = torch.matmul(Q / dk ** 0.5, K.transpose(2, 3)).masked_fill(pair == 0, -1e10)
        att = F.softmax(att, dim=-1)
        if training and attn_dropout > 0:
            att = F.dropout(att, attn_dropout, True)
        att = att * pair
        o = torch.matmul(att, V).transpose(1, 2).contiguous().view(N, k, -1)
        o = F.linear(o, sd[q + "slf_attn.fc.weight"]) + res
        x = _masked_ln(o, mask, sd, q + "slf_attn.norm.") * mk
        res = x
        h = _relu(F.linear(x, sd[q + "pos_ffn.w_1.weight"], sd[q + "pos_ffn.w_1.bias"])) * mk
        h = F.linear(h, sd[q + "pos_ffn.w_2.weight"], sd[q + "pos_ffn.w_2.bias"]) * mk
        x = _masked_ln(h + res, mask, sd, q + "pos_ffn.norm.") * mk
    x = x.sum(dim=1)
    x = F.linear(x, sd[p + "out.0.weight"])
    return _bn(x, sd, p + "out.1.", training)


def sign_net(eigS, eigV, edge_index, batch, sd, p, nl_phi, nl_rho, ignore_eigval=False, training=True,
             attn_dropout=0.0):
    """SignNet.forward (sign_net.py:96-118) from the dense-list tensors: [N,k] -> [N,n_hid]."""
    mask = slot_mask(batch, eigV.shape[1])
    pos = 0
    if not ignore_eigval:
        pos = masked_mlp(eigS.unsqueeze(-1), mask, sd, p + "eigen_encoder.", 2, True, training)
    x = phi_pm(eigV, edge_index, mask, sd, p + "phi.", nl_phi, training)
    return set_transformer(x, pos, mask, sd, p + "rho.", nl_rho, 4, training, attn_dropout)


# --------------------------------------------------------------------------------------------------------------------
# PyG flavour: downstream GNN predictor with GINEConv (row a12)
# --------------------------------------------------------------------------------------------------------------------
def _mlp(x, sd, p, nlayer, with_final_activation=True, with_norm=True, training=True):
    """elements.MLP.forward (elements.py:39-69)."""
    for i in range(nlayer):
        x = F.linear(x, sd[f"{p}layers.{i}.weight"], sd.get(f"{p}layers.{i}.bias"))
        if i < nlayer - 1 or with_final_activation:
            if with_norm:
                x = _bn(x, sd, f"{p}norms.{i}.", training)
            x = _relu(x)
    return x


def _discrete_encoder(x, sd, p):
    """DiscreteEncoder.forward (elements.py:31-37): sum of per-column embeddings."""
    if x.dim() == 1:
        x = x.unsqueeze(1)
    out = 0
    for i in range(x.size(1)):
        out = out + F.embedding(x[:, i], sd[f"{p}embeddings.{i}.weight"])
    return out


def gine_aggregate(x, edge_index, edge_emb, eps):
    """K6.  PyG GINEConv core: sum_j relu(x_j + e_ij) into i, then += (1+eps) x_i (pyg_gnn_wrapper.py:23,28)."""
    msg = (x.index_select(0, edge_index[0]) + edge_emb).relu()
    out = torch.zeros_like(x).index_add_(0, edge_index[1], msg)
    return out + (1 + eps) * x


def gnn_predictor(x_in, edge_index, edge_attr, batch, pos, sd, p, nlayer, num_graphs=None, training=True):
    """GNN.forward (Alchemy/sign_net/model.py:36-64), gnn_type='GINEConv', pooling='add', bn, res.
    Discrete (ZINC) vs continuous (Alchemy) inputs are told apart by the state_dict keys, as in model.py:13-14."""
    if f"{p}input_encoder.embeddings.0.weight" in sd:
        x = _discrete_encoder(x_in.squeeze(), sd, p + "input_encoder.")
    else:
        x = _mlp(x_in.squeeze(), sd, p + "input_encoder.", 1, training=training)
    if pos is not None:
        x = F.linear(torch.cat([x, pos], dim=-1), sd[p + "linear.weight"], sd[p + "linear.bias"])
    if edge_attr is None:
        edge_attr = edge_index.new_zeros(edge_index.size(-1))
    prev = x
    for l in range(nlayer):
        if f"{p}edge_encoders.{l}.embeddings.0.weight" in sd:
            e = _discrete_encoder(edge_attr, sd, f"{p}edge_encoders.{l}.")
        else:
            e = _mlp(edge_attr, sd, f"{p}edge_encoders.{l}.", 1, training=training)
        a = gine_aggregate(x, edge_index, e, sd[f"{p}convs.{l}.layer.eps"])
        x = _mlp(a, sd, f"{p}convs.{l}.nn.", 2, False, True, training)
        x = _relu(_bn(x, sd, f"{p}norms.{l}.", training))
        x = x + prev
        prev = x
    B = int(batch.max()) + 1 if num_graphs is None else num_graphs
    x = x.new_zeros(B, x.shape[1]).index_add_(0, batch, x)  # K5 add-pool (model.py:61)
    return _mlp(x, sd, p + "output_encoder.", 2, False, True, training)


def sign_net_gnn(data, sd, nl_signnet, nl_gnn, nl_rho=4, ignore_eigval=False, training=True, attn_dropout=0.0):
    """SignNetGNN.forward (sign_net.py:130-132)."""
    eigS, eigV = dense_list_evd(data.eigen_values, data.eigen_vectors, data.batch)
    pos = sign_net(eigS, eigV, data.edge_index, data.batch, sd, "sign_net.", nl_signnet, nl_rho, ignore_eigval, training,
                   attn_dropout)
    return gnn_predictor(data.x, data.edge_index, data.edge_attr, data.batch, pos, sd, "gnn.", nl_gnn, training=training)


# --------------------------------------------------------------------------------------------------------------------
# DGL flavour (rows a9-a11): GIN phi on [N,k,C] with unmasked BN, rho = MLP
# --------------------------------------------------------------------------------------------------------------------
def _bn_nkc(x, sd, p, training):
    """BatchNorm1d applied as bn(x.transpose(2,1)).transpose(2,1) on [N,k,C] == BN over all N*k rows (padded slots
    included), gnns.py:107-112 / mlp.py:42-46."""
    if x.ndim == 2:
        return _bn(x, sd, p, training)
    sh = x.shape
    return _bn(x.reshape(-1, sh[-1]), sd, p, training).reshape(sh)


def dgl_mlp(x, sd, p, num_layers, training=True):
    """layers/mlp.py:37-56 with use_bn=True, relu, dropout 0: (Linear -> ReLU -> BN) x (L-1) -> Linear."""
    for i in range(num_layers - 1):
        x = _relu(F.linear(x, sd[f"{p}lins.{i}.weight"], sd[f"{p}lins.{i}.bias"]))
        x = _bn_nkc(x, sd, f"{p}bns.{i}.", training)
    i = num_layers - 1
    return F.linear(x, sd[f"{p}lins.{i}.weight"], sd[f"{p}lins.{i}.bias"])


def dgl_gin(x, src, dst, sd, p, n_layers, training=True):
    """GIN.forward (layers/gnns.py:102-114): layer i>0 is preceded by BN_{i-1}; dgl GINConv('sum', eps buffer 0):
    rst = (1+eps)*feat + sum_{u->v} feat_u ; apply_func = 2-layer MLP."""
    for i in range(n_layers):
        if i != 0:
            x = _bn_nkc(x, sd, f"{p}bns.{i - 1}.", training)
        eps = sd.get(f"{p}layers.{i}.eps", torch.zeros(1))
        neigh = torch.zeros_like(x).index_add_(0, dst, x.index_select(0, src))
        x = dgl_mlp((1 + eps) * x + neigh, sd, f"{p}layers.{i}.apply_func.", 2, training)
    return x


def gin_deepsigns(x, src, dst, sd, n_layers, k, training=True):
    """GINDeepSigns.forward (layers/deepsigns.py:45-51): x [N,k,1] -> [N,k,1]."""
    h = dgl_gin(x, src, dst, sd, "enc.", n_layers, training) + dgl_gin(-x, src, dst, sd, "enc.", n_layers, training)
    h = dgl_mlp(h.reshape(h.shape[0], -1), sd, "rho.", n_layers, training)
    return h.reshape(x.shape[0], k, 1)


def masked_gin_deepsigns(x, src, dst, num_nodes_per_graph, sd, n_layers, k, training=True):
    """MaskedGINDeepSigns.forward (layers/deepsigns.py:72-86): zero slots >= n_b, sum over k, rho MLP."""
    h = dgl_gin(x, src, dst, sd, "enc.", n_layers, training) + dgl_gin(-x, src, dst, sd, "enc.", n_layers, training)
    n_per_node = torch.repeat_interleave(num_nodes_per_graph, num_nodes_per_graph)
    mask = torch.arange(x.shape[1])[None, :] < n_per_node[:, None]
    h = (h * mask.unsqueeze(-1)).sum(dim=1)
    h = dgl_mlp(h, sd, "rho.", n_layers, training)
    return h.reshape(x.shape[0], k, 1)


def gin_net(h_idx, pos_enc, src, dst, num_nodes_per_graph, sd, n_layers, readout="mean", training=True, p=""):
    """GINNet.forward, `pe_init='lap_pe'`, no LSPE (GraphPrediction/nets/ZINC_graph_regression/gin_net.py:81-138):
    h = embedding_h(atom) + embedding_p(pos_enc); L x dgl GINConv(MLP(2 layers), 'sum'); mean/sum readout; MLPReadout
    (layers/mlp_readout_layer.py:9-25).  embedding_e is computed and discarded by the reference (GINConv takes no e)."""
    h = sd[p + "embedding_h.weight"][h_idx] + F.linear(pos_enc, sd[p + "embedding_p.weight"], sd[p + "embedding_p.bias"])
    for l in range(n_layers):
        eps = sd.get(f"{p}layers.{l}.eps", torch.zeros(1))
        rst = (1 + eps) * h + torch.zeros_like(h).index_add_(0, dst, h.index_select(0, src))
        h = dgl_mlp(rst, sd, f"{p}layers.{l}.apply_func.", 2, training)
    n = torch.as_tensor(num_nodes_per_graph)
    seg = torch.repeat_interleave(torch.arange(n.numel()), n)
    hg = torch.zeros(n.numel(), h.shape[1], dtype=h.dtype).index_add_(0, seg, h)
    if readout != "sum":
        hg = hg / n.to(h.dtype).clamp(min=1).unsqueeze(1)
    y, L = hg, 2
    for l in range(L):
        y = _relu(F.linear(y, sd[f"{p}MLP_layer.FC_layers.{l}.weight"], sd[f"{p}MLP_layer.FC_layers.{l}.bias"]))
    return F.linear(y, sd[f"{p}MLP_layer.FC_layers.{L}.weight"], sd[f"{p}MLP_layer.FC_layers.{L}.bias"])


def gated_gcn_layer(h, e, src, dst, sd, p, batch_norm=True, residual=True, training=True):
    """GatedGCNLayer.forward with graph_norm=False, dropout 0 (GraphPrediction/layers/gatedgcn_layer.py:36-77):
    e' = D h_src + E h_dst + C e;  sigma = sigmoid(e');  h' = A h + (sum_in B h_src * sigma) / (sum_in sigma + 1e-6);
    BatchNorm1d on both, ReLU, residual (dropped when the layer changes the width, :24-25)."""
    lin = lambda x, n: F.linear(x, sd[f"{p}{n}.weight"], sd[f"{p}{n}.bias"])
    Ah, Bh, Dh, Eh, Ce = lin(h, "A"), lin(h, "B"), lin(h, "D"), lin(h, "E"), lin(e, "C")
    e_new = (Dh.index_select(0, src) + Eh.index_select(0, dst)) + Ce
    sigma = torch.sigmoid(e_new)
    N = h.shape[0]
    ssh = torch.zeros(N, Bh.shape[1], dtype=h.dtype).index_add(0, dst, Bh.index_select(0, src) * sigma)
    ss = torch.zeros(N, Bh.shape[1], dtype=h.dtype).index_add(0, dst, sigma)
    h_new = Ah + ssh / (ss + 1e-6)
    if batch_norm:
        h_new = _bn(h_new, sd, f"{p}bn_node_h.", training)
        e_new = _bn(e_new, sd, f"{p}bn_node_e.", training)
    h_new, e_new = _relu(h_new), _relu(e_new)
    if residual and h.shape[1] == h_new.shape[1]:
        h_new, e_new = h + h_new, e + e_new
    return h_new, e_new


def gatedgcn_net(h_idx, pos_enc, e_idx, src, dst, num_nodes_per_graph, sd, n_layers, readout="mean", edge_feat=True,
                 pe_aggregate="add", batch_norm=True, residual=True, training=True, p=""):
    """GatedGCNNet.forward, `pe_init='lap_pe'`, no LSPE (GraphPrediction/nets/ZINC_graph_regression/gatedgcn_net.py:86-135):
    h = embedding_h(atom) (+ | concat-project) embedding_p(pos_enc); e = embedding_e(bond) (or Linear(1) of ones);
    L x GatedGCNLayer; mean/sum readout; MLPReadout (layers/mlp_readout_layer.py:9-25)."""
    h = sd[p + "embedding_h.weight"][h_idx]
    pp = F.linear(pos_enc, sd[p + "embedding_p.weight"], sd[p + "embedding_p.bias"])
    if pe_aggregate == "concat":
        h = F.linear(torch.cat([h, pp], dim=1), sd[p + "pe_proj.weight"], sd[p + "pe_proj.bias"])
    else:
        h = h + pp
    if edge_feat:
        e = sd[p + "embedding_e.weight"][e_idx]
    else:
        e = F.linear(torch.ones(src.numel(), 1, dtype=h.dtype), sd[p + "embedding_e.weight"], sd[p + "embedding_e.bias"])
    for l in range(n_layers):
        h, e = gated_gcn_layer(h, e, src, dst, sd, f"{p}layers.{l}.", batch_norm, residual, training)
    n = torch.as_tensor(num_nodes_per_graph)
    seg = torch.repeat_interleave(torch.arange(n.numel()), n)
    hg = torch.zeros(n.numel(), h.shape[1], dtype=h.dtype).index_add_(0, seg, h)
    if readout != "sum":
        hg = hg / n.to(h.dtype).clamp(min=1).unsqueeze(1)
    y, L = hg, 2
    for l in range(L):
        y = _relu(F.linear(y, sd[f"{p}MLP_layer.FC_layers.{l}.weight"], sd[f"{p}MLP_layer.FC_layers.{l}.bias"]))
    return F.linear(y, sd[f"{p}MLP_layer.FC_layers.{L}.weight"], sd[f"{p}MLP_layer.FC_layers.{L}.bias"])


def pna_aggregate(msg, dst, N, avg_d_log, aggregators=("mean", "max", "min", "std"),
                  scalers=("identity", "amplification", "attenuation")):
    """PNATower.reduce_func_for_h (GraphPrediction/layers/pna_layer.py:50-55, pna_utils.py:12-31,73-84): per destination
    node, over its D incoming messages: cat over aggregators [mean | max | min | sqrt(relu(E[x^2] - E[x]^2) + 1e-5)],
    then cat over scalers [identity | * log(D+1)/avg_d | * avg_d/log(D+1)].  Nodes without incoming edges keep zeros
    (DGL does not call the reduce function for them)."""
    C = msg.shape[1]
    deg = torch.bincount(dst, minlength=N)
    D = deg.clamp(min=1).to(msg.dtype).unsqueeze(1)
    idx = dst.unsqueeze(1).expand(-1, C)
    s1 = torch.zeros(N, C, dtype=msg.dtype).index_add(0, dst, msg)
    s2 = torch.zeros(N, C, dtype=msg.dtype).index_add(0, dst, msg * msg)
    mean = s1 / D
    outs = {"mean": mean,
            "max": torch.full((N, C), float("-inf"), dtype=msg.dtype).scatter_reduce(0, idx, msg, "amax", include_self=True),
            "min": torch.full((N, C), float("inf"), dtype=msg.dtype).scatter_reduce(0, idx, msg, "amin", include_self=True),
            "std": torch.sqrt(torch.relu(s2 / D - mean * mean) + 1e-5), "sum": s1}
    h = torch.cat([outs[a] for a in aggregators], dim=1)
    logd = torch.log(deg.clamp(min=1).to(msg.dtype) + 1).unsqueeze(1)
    scale = {"identity": torch.ones_like(logd), "amplification": logd / avg_d_log, "attenuation": avg_d_log / logd}
    h = torch.cat([h * scale[sc] for sc in scalers], dim=1)
    return torch.where((deg > 0).unsqueeze(1), h, torch.zeros_like(h))


def pna_layer(h, e, src, dst, snorm_n, sd, p, towers, divide_input, avg_d_log, graph_norm=True, batch_norm=True,
              residual=True, edge_features=True, training=True, aggregators=("mean", "max", "min", "std"),
              scalers=("identity", "amplification", "attenuation")):
    """PNALayer.forward with pretrans_layers = posttrans_layers = 1, dropout 0 (pna_layer.py:16-82,134-153): per tower
    Linear(cat[h_src, h_dst, e]) -> aggregate -> Linear(cat[h, agg]) -> * snorm_n -> BatchNorm; towers concatenated ->
    Linear + LeakyReLU(0.01) mixing network -> residual (dropped when the layer changes the width)."""
    N, in_dim = h.shape
    tin = in_dim // towers if divide_input else in_dim
    outs = []
    for t in range(towers):
        ht = h[:, t * tin:(t + 1) * tin] if divide_input else h
        q = f"{p}towers.{t}."
        z = [ht.index_select(0, src), ht.index_select(0, dst)] + ([e] if edge_features else [])
        msg = F.linear(torch.cat(z, dim=1), sd[q + "pretrans_h.fully_connected.0.linear.weight"],
                       sd[q + "pretrans_h.fully_connected.0.linear.bias"])
        agg = pna_aggregate(msg, dst, N, avg_d_log, aggregators, scalers)
        ho = F.linear(torch.cat([ht, agg], dim=1), sd[q + "posttrans_h.fully_connected.0.linear.weight"],
                      sd[q + "posttrans_h.fully_connected.0.linear.bias"])
        if graph_norm:
            ho = ho * snorm_n
        if batch_norm:
            ho = _bn(ho, sd, q + "batchnorm_h.", training)
        outs.append(ho)
    hc = torch.cat(outs, dim=1)
    ho = F.leaky_relu(F.linear(hc, sd[p + "mixing_network_h.linear.weight"], sd[p + "mixing_network_h.linear.bias"]), 0.01)
    if residual and ho.shape[1] == in_dim:
        ho = h + ho
    return ho


def pna_net(h_idx, pos_enc, e_idx, src, dst, num_nodes_per_graph, snorm_n, sd, n_layers, towers, avg_d_log,
            readout="sum", divide_input_first=True, divide_input_last=True, graph_norm=True, batch_norm=True,
            residual=True, edge_feat=True, training=True, p=""):
    """PNANet.forward, `pe_init='lap_pe'`, no LSPE, gru=False (GraphPrediction/nets/ZINC_graph_regression/pna_net.py:116-167):
    h = embedding_h(atom) + embedding_p(pos_enc); e = embedding_e(bond); L x PNALayer; readout; MLPReadout."""
    h = sd[p + "embedding_h.weight"][h_idx] + F.linear(pos_enc, sd[p + "embedding_p.weight"], sd[p + "embedding_p.bias"])
    e = sd[p + "embedding_e.weight"][e_idx] if edge_feat else None
    for l in range(n_layers):
        div = divide_input_first if l < n_layers - 1 else divide_input_last
        h = pna_layer(h, e, src, dst, snorm_n, sd, f"{p}layers.{l}.", towers, div, avg_d_log, graph_norm, batch_norm,
                      residual, edge_feat, training)
    n = torch.as_tensor(num_nodes_per_graph)
    seg = torch.repeat_interleave(torch.arange(n.numel()), n)
    hg = torch.zeros(n.numel(), h.shape[1], dtype=h.dtype).index_add_(0, seg, h)
    if readout != "sum":
        hg = hg / n.to(h.dtype).clamp(min=1).unsqueeze(1)
    y, L = hg, 2
    for l in range(L):
        y = _relu(F.linear(y, sd[f"{p}MLP_layer.FC_layers.{l}.weight"], sd[f"{p}MLP_layer.FC_layers.{l}.bias"]))
    return F.linear(y, sd[f"{p}MLP_layer.FC_layers.{L}.weight"], sd[f"{p}MLP_layer.FC_layers.{L}.bias"])


def sparse_attention(Qh, Kh, Eh, Vh, src, dst, n_heads):
    """MultiHeadAttentionLayer.propagate_attention with full_graph=False (GraphPrediction/layers/transformer.py:160-192):
    per edge k: j -> i and head: a = sum_c K[j] Q[i] E[k] / sqrt(d);  s = exp(clamp(a, -5, 5));
    out_i = (sum_in s V[j]) / (sum_in s + 1e-6).  [N or E, H*d] in, [N, H*d] out."""
    N, HD = Qh.shape
    d = HD // n_heads
    v = lambda t: t.reshape(-1, n_heads, d)
    score = (v(Kh).index_select(0, src) * v(Qh).index_select(0, dst)) / (d ** 0.5)
    score = score * v(Eh)
    s = torch.exp(score.sum(-1, keepdim=True).clamp(-5, 5))
    wV = torch.zeros(N, n_heads, d, dtype=Qh.dtype).index_add(0, dst, v(Vh).index_select(0, src) * s)
    z = torch.zeros(N, n_heads, 1, dtype=Qh.dtype).index_add(0, dst, s)
    return (wV / (z + 1e-6)).reshape(N, HD)


def transformer_layer(h, e, src, dst, sd, p, n_heads, batch_norm=True, residual=True, training=True):
    """BatchedTransformerLayer.forward as TransformerNet builds it (transformer.py:232-301; transformer_net.py:68-69 passes
    neither layer_norm nor use_bias, so: no LayerNorm, bias-free Q/K/E/V): attention -> O_h -> residual -> BatchNorm ->
    FFN (Linear, ReLU, Linear) -> residual -> BatchNorm."""
    lin = lambda x, n: F.linear(x, sd[f"{p}{n}.weight"], sd.get(f"{p}{n}.bias"))
    a = sparse_attention(lin(h, "attention_h.Q"), lin(h, "attention_h.K"), lin(e, "attention_h.E"), lin(h, "attention_h.V"),
                         src, dst, n_heads)
    x = lin(a, "O_h")
    if residual:
        x = h + x
    if batch_norm:
        x = _bn(x, sd, f"{p}batch_norm1_h.", training)
    y = lin(_relu(lin(x, "FFN_h_layer1")), "FFN_h_layer2")
    if residual:
        y = x + y
    if batch_norm:
        y = _bn(y, sd, f"{p}batch_norm2_h.", training)
    return y


def transformer_net(h_idx, pos_enc, e_idx, src, dst, num_nodes_per_graph, sd, n_layers, n_heads, readout="sum",
                    pe_aggregate="concat", batch_norm=True, residual=True, training=True, p=""):
    """TransformerNet.forward, `pe_init='lap_pe'`, no LSPE, full_graph=False, edge_feat=True
    (GraphPrediction/nets/ZINC_graph_regression/transformer_net.py:89-150)."""
    h = sd[p + "embedding_h.weight"][h_idx]
    pp = F.linear(pos_enc, sd[p + "embedding_p.weight"], sd[p + "embedding_p.bias"])
    if pe_aggregate == "concat":
        h = F.linear(torch.cat([h, pp], dim=1), sd[p + "pe_proj.weight"], sd[p + "pe_proj.bias"])
    else:
        h = h + pp
    e = sd[p + "embedding_e.weight"][e_idx]
    for l in range(n_layers):
        h = transformer_layer(h, e, src, dst, sd, f"{p}layers.{l}.", n_heads, batch_norm, residual, training)
    n = torch.as_tensor(num_nodes_per_graph)
    seg = torch.repeat_interleave(torch.arange(n.numel()), n)
    hg = torch.zeros(n.numel(), h.shape[1], dtype=h.dtype).index_add_(0, seg, h)
    if readout != "sum":
        hg = hg / n.to(h.dtype).clamp(min=1).unsqueeze(1)
    y, L = hg, 2
    for l in range(L):
        y = _relu(F.linear(y, sd[f"{p}MLP_layer.FC_layers.{l}.weight"], sd[f"{p}MLP_layer.FC_layers.{l}.bias"]))
    return F.linear(y, sd[f"{p}MLP_layer.FC_layers.{L}.weight"], sd[f"{p}MLP_layer.FC_layers.{L}.bias"])


def handle_lap(pos_enc, num_nodes_per_graph, lap_method, sign_flip=None):
    """The positional-encoding baselines of train/train_ZINC_graph_regression.py:12-47 other than `sign_inv`:
    'sign_flip' (random column signs; the caller passes the draw), 'abs_val', 'canonical' (per graph and column:
    flip when the column has fewer non-negative than negative entries OR less non-negative than negative mass;
    `less_nonneg + less_norm` is a boolean OR in torch), 'none'."""
    if lap_method == "sign_flip":
        return pos_enc * sign_flip.unsqueeze(0)
    if lap_method == "abs_val":
        return pos_enc.abs()
    if lap_method == "none":
        return pos_enc
    if lap_method != "canonical":
        raise ValueError("invalid laplacian method")
    n = torch.as_tensor(num_nodes_per_graph)
    seg = torch.repeat_interleave(torch.arange(n.numel()), n)
    seg_sum = lambda x: torch.zeros(n.numel(), x.shape[1], dtype=x.dtype).index_add_(0, seg, x)
    less_nonneg = seg_sum((pos_enc >= 0).float()) < seg_sum((pos_enc < 0).float())
    nonneg, neg = pos_enc.clone(), pos_enc.clone()
    nonneg[pos_enc < 0] = 0
    neg[pos_enc >= 0] = 0
    less_norm = seg_sum(nonneg) < seg_sum(neg.abs())
    flip = -(less_nonneg + less_norm).float()
    flip[flip == 0] = 1
    return flip.index_select(0, seg) * pos_enc


# --------------------------------------------------------------------------------------------------------------------
# LearningFilters flavour (rows a14, a15): single-graph SignNet (DeepSets phi) and BasisNet IGN 2->1
# --------------------------------------------------------------------------------------------------------------------
def eq_deepsets(x, sd, p, num_layers, use_bn=True):
    """EqDeepSetsEncoder.forward (LearningFilters/models.py:91-113): per-element Linear + Linear(mean over the set
    dim -2); ReLU; BN(track_running_stats=False => always batch stats)."""
    for i in range(num_layers - 1):
        x = _relu(F.linear(x, sd[f"{p}lins1.{i}.weight"], sd[f"{p}lins1.{i}.bias"])
                   + F.linear(x.mean(dim=-2, keepdim=True), sd[f"{p}lins2.{i}.weight"], sd[f"{p}lins2.{i}.bias"]))
        if use_bn:
            sh = x.shape
            x = F.batch_norm(x.reshape(-1, sh[-1]), None, None, sd[f"{p}bns.{i}.weight"], sd[f"{p}bns.{i}.bias"],
                             True, BN_MOMENTUM, BN_EPS).reshape(sh)
    i = num_layers - 1
    return (F.linear(x, sd[f"{p}lins1.{i}.weight"], sd[f"{p}lins1.{i}.bias"])
            + F.linear(x.mean(dim=-2, keepdim=True), sd[f"{p}lins2.{i}.weight"], sd[f"{p}lins2.{i}.bias"]))


def sign_plus_deepsets(x, sd, p, num_layers, use_bn=True):
    """SignPlus.forward (LearningFilters/signbasisnet.py:16-20): model(x) + model(-x)."""
    return eq_deepsets(x, sd, p, num_layers, use_bn) + eq_deepsets(-x, sd, p, num_layers, use_bn)


def ign_2to1_ops(P):
    """contractions_2_to_1 (LearningFilters/ign.py:344-374) on P [b, d, m, m] -> [b, d, 5, m]:
    diag, trace/m, rowsum/m, colsum/m, total/m^2 (normalised)."""
    m = P.shape[-1]
    diag = torch.diagonal(P, dim1=-2, dim2=-1)
    tr = diag.sum(-1, keepdim=True)
    row = P.sum(dim=3)
    col = P.sum(dim=2)
    tot = row.sum(-1, keepdim=True)
    return torch.stack([diag, tr.expand_as(diag) / m, row / m, col / m, tot.expand_as(diag) / m ** 2], dim=2)


def ign_1to1_ops(x):
    """contractions_1_to_1 (ign.py:404-417): identity and mean over the set."""
    return torch.stack([x, x.sum(dim=2, keepdim=True).expand_as(x) / x.shape[2]], dim=2)


def ign2to1(P, sd, p="", training=True):
    """IGN2to1.forward (LearningFilters/ign.py:29-39): 2->1 equivariant layer, two 1->1 layers (each: einsum with
    coeffs [D,S,basis] + bias, ReLU, then BN over channel dim of [b,S,m]), then fc1/ReLU/fc2 on the channel dim.
    P [b,1,m,m] -> [b,out,m]."""
    def bn3(x, q):
        y = F.batch_norm(x, sd[q + "running_mean"], sd[q + "running_var"], sd[q + "weight"], sd[q + "bias"],
                         training, BN_MOMENTUM, BN_EPS)
        if training and (q + "num_batches_tracked") in sd:
            sd[q + "num_batches_tracked"] += 1
        return y

    ops = [ign_2to1_ops, ign_1to1_ops, ign_1to1_ops]
    x = P
    for i in range(3):
        x = torch.einsum("dsb,ndbi->nsi", sd[f"{p}equi_layers.{i}.coeffs"], ops[i](x)) + sd[f"{p}equi_layers.{i}.bias"]
        x = bn3(_relu(x), f"{p}bns.{i}.")
    x = x.transpose(2, 1)
    x = _relu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    x = F.linear(x, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
    return x.transpose(2, 1)


def eigenspace_groups(eigvals: torch.Tensor, decimals: int = 5):
    """Eigenvalue grouping of LearningFilters/training.py:47-62: round to `decimals`, unique -> (counts, sections).
    Returns a list of (start, stop) column ranges, one per eigenspace, in ascending eigenvalue order."""
    r = torch.round(eigvals * 10 ** decimals) / (10 ** decimals)
    _, counts = r.unique(return_counts=True)
    stops = torch.cumsum(counts, 0).tolist()
    starts = [0] + stops[:-1]
    return list(zip(starts, stops))

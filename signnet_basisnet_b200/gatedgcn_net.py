"""DGL-flavour GatedGCN predictor consuming the sign-invariant positional encoding — the base
model of `configs/gatedgcn/GatedGCN_ZINC_LapPE_signinv_GIN_mask.json` (k = 37, the configuration cfg 4 is quoted on) —
and the non-learned PE baselines of the training loop.

Mirrors GraphPrediction/nets/ZINC_graph_regression/gatedgcn_net.py:19-135 (`GatedGCNNet`, the `pe_init='lap_pe'` /
no-LSPE path: the LSPE layer class is not even importable in the reference file) and layers/gatedgcn_layer.py:12-77
(`GatedGCNLayer`): same `net_params` keys, same state_dict keys (`embedding_h`, `embedding_p`, `embedding_e`, `pe_proj`,
`layers.{l}.{A,B,C,D,E}.{weight,bias}`, `layers.{l}.bn_node_{h,e}.*`, `MLP_layer.FC_layers.*`, `sign_inv_net.*`),
`forward(g, h, p, e, snorm_n) -> (scores, g)` as called at train/train_ZINC_graph_regression.py:76, and
`handle_lap(model, batch_pos_enc, batch_graphs, device)` (:12-47).

Kernels: the five Linears of a layer = sb_linear_fwd / sb_linear_wgrad (tcgen05 above the small-problem threshold), the
dgl message passing (apply_edges(u_add_v) + two update_all) = ONE fused edge-gated aggregate sb_gated_agg_fwd/bwd
(csrc/gated.cu), BatchNorm + ReLU + residual of both streams = the BatchNorm kernels of phi, read-out = sb_segment_pool.
STATUS: parity pinned on the CPU side (oracle/restate.gatedgcn_net vs the reference class, golden fixture
tests/golden/dgl_gatedgcn_net.pt); GPU parity (kernel vs oracle, GatedGCNNet vs the reference fixture,
PE baselines): tests/test_gpu_gatedgcn.py, green on the B200.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import counted_call as _call, ptr as _p
from .deepsigns import _graph_index, get_sign_inv_net
from .functional import add_rows, batch_norm_act, linear
from .gin_net import MLPReadout
from .layout import pad4
from .model import EmbeddingSumFn, Linear2Fn, SegmentPoolFn


class GatedAggFn(torch.autograd.Function):
    """(Ah, Bh, Dh, Eh [N, ld], Ce [E, ld]) -> (h' [N, ld], e' [E, ld]); gatedgcn_layer.py:48-54."""

    @staticmethod
    def forward(ctx, Ah, Bh, Dh, Eh, Ce, gi):
        Ah, Bh, Dh, Eh, Ce = (t.contiguous() for t in (Ah, Bh, Dh, Eh, Ce))
        N, ld = Ah.shape
        E = Ce.shape[0]
        if E != gi.E or N != gi.N or Ce.shape[1] != ld:
            raise ValueError("gated aggregate: node / edge tensors do not match the graph")
        dev = Ah.device
        h = torch.empty(N, ld, dtype=torch.float32, device=dev)
        e = torch.empty(E, ld, dtype=torch.float32, device=dev)
        ss = torch.empty(N, ld, dtype=torch.float32, device=dev)
        ssh = torch.empty(N, ld, dtype=torch.float32, device=dev)
        _call("sb_gated_agg_fwd", _p(Ah), _p(Bh), _p(Dh), _p(Eh), _p(Ce), _p(gi.in_ptr), _p(gi.in_src), _p(gi.in_eid), N,
              ld, _p(e), _p(h), _p(ss), _p(ssh))
        ctx.save_for_backward(Bh, e, ss, ssh)
        ctx.gi = gi
        ctx.set_materialize_grads(False)   # the last layer's edge stream has no consumer: de arrives as None
        return h, e

    @staticmethod
    def backward(ctx, dh, de):
        Bh, e, ss, ssh = ctx.saved_tensors
        gi = ctx.gi
        N, ld = Bh.shape
        dev = Bh.device
        dh = torch.zeros(N, ld, dtype=torch.float32, device=dev) if dh is None else dh.contiguous()
        de = None if de is None else de.contiguous()
        dBh, dDh, dEh = (torch.empty(N, ld, dtype=torch.float32, device=dev) for _ in range(3))
        dCe = torch.empty(gi.E, ld, dtype=torch.float32, device=dev)
        _call("sb_gated_agg_bwd", _p(dh), _p(de), _p(Bh), _p(e), _p(ss), _p(ssh), _p(gi.edge_index), _p(gi.in_ptr),
              _p(gi.in_eid), _p(gi.out_ptr), _p(gi.out_dst), _p(gi.out_eid), N, gi.E, ld, _p(dBh), _p(dDh), _p(dEh),
              _p(dCe))
        return dh, dBh, dDh, dEh, dCe, None


class GatedGCNLayer(nn.Module):
    """layers/gatedgcn_layer.py:12-77 with dropout 0 (every shipped sign_inv configuration)."""

    def __init__(self, input_dim, output_dim, dropout, batch_norm, residual=False, graph_norm=True):
        super().__init__()
        if dropout != 0:
            raise NotImplementedError("dropout is 0.0 in every shipped sign_inv configuration")
        self.in_channels, self.out_channels = input_dim, output_dim
        self.batch_norm, self.graph_norm = batch_norm, graph_norm
        self.residual = residual and input_dim == output_dim       # gatedgcn_layer.py:24-25
        self.A = nn.Linear(input_dim, output_dim, bias=True)
        self.B = nn.Linear(input_dim, output_dim, bias=True)
        self.C = nn.Linear(input_dim, output_dim, bias=True)
        self.D = nn.Linear(input_dim, output_dim, bias=True)
        self.E = nn.Linear(input_dim, output_dim, bias=True)
        self.bn_node_h = nn.BatchNorm1d(output_dim)
        self.bn_node_e = nn.BatchNorm1d(output_dim)

    def forward_rows(self, gi, h, e, snorm_n=None):
        """h [N, pad4(in)], e [E, pad4(in)] -> (h, e) padded to pad4(out)."""
        if self.graph_norm:
            raise NotImplementedError("graph_norm=True is never selected by GatedGCNNet (gatedgcn_net.py:67-69)")
        if not self.batch_norm:
            raise NotImplementedError("batch_norm=False is not built (no shipped configuration selects it)")
        ld = pad4(self.out_channels)
        lin = lambda x, m: linear(x, m.weight, m.bias, ld)
        h_new, e_new = GatedAggFn.apply(lin(h, self.A), lin(h, self.B), lin(h, self.D), lin(h, self.E), lin(e, self.C), gi)
        h_new = batch_norm_act(h_new, self.bn_node_h, self.training, relu=True, res=h if self.residual else None)
        e_new = batch_norm_act(e_new, self.bn_node_e, self.training, relu=True, res=e if self.residual else None)
        return h_new, e_new

    def forward(self, g, h, p=None, e=None, snorm_n=None):
        gi = _graph_index(g, h.device)
        h, e = self.forward_rows(gi, h, e, snorm_n)
        return h[:, :self.out_channels], None, e[:, :self.out_channels]

    def __repr__(self):
        return "{}(in_channels={}, out_channels={})".format(self.__class__.__name__, self.in_channels, self.out_channels)


class GatedGCNNet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        hidden_dim, out_dim = net_params["hidden_dim"], net_params["out_dim"]
        self.n_layers = net_params["L"]
        self.readout = net_params["readout"]
        self.batch_norm = net_params["batch_norm"]
        self.residual = net_params["residual"]
        self.edge_feat = net_params["edge_feat"]
        self.device = net_params["device"]
        self.pe_init = net_params["pe_init"]
        self.lap_method = net_params["lap_method"]
        self.lap_lspe = net_params["lap_lspe"]
        self.use_lapeig_loss = net_params["use_lapeig_loss"]
        self.lambda_loss = net_params["lambda_loss"]
        self.alpha_loss = net_params["alpha_loss"]
        self.pos_enc_dim = net_params["pos_enc_dim"]
        self.pe_aggregate = net_params["pe_aggregate"]
        if self.pe_init == "rand_walk" or self.lap_lspe:
            raise NotImplementedError("GatedGCNNet on the B200 path: the LSPE branch is out of scope (the reference "
                                      "file does not import its layer class either)")
        if self.pe_init != "lap_pe":
            raise NotImplementedError("GatedGCNNet on the B200 path is the `pe_init='lap_pe'` predictor")
        if net_params["in_feat_dropout"] != 0 or net_params["dropout"] != 0:
            raise NotImplementedError("dropout is 0.0 in every shipped sign_inv configuration")
        if self.use_lapeig_loss:
            raise NotImplementedError("the Laplacian-eigenvector loss belongs to the LSPE branch (out of scope)")
        self.embedding_p = nn.Linear(self.pos_enc_dim, hidden_dim)
        self.embedding_h = nn.Embedding(net_params["num_atom_type"], hidden_dim)
        self.embedding_e = (nn.Embedding(net_params["num_bond_type"], hidden_dim) if self.edge_feat
                            else nn.Linear(1, hidden_dim))
        mk = lambda o: GatedGCNLayer(hidden_dim, o, 0.0, self.batch_norm, residual=self.residual, graph_norm=False)
        self.layers = nn.ModuleList([mk(hidden_dim) for _ in range(self.n_layers - 1)] + [mk(out_dim)])
        self.MLP_layer = MLPReadout(out_dim, 1)
        self.hidden_dim, self.out_dim = hidden_dim, out_dim
        self.g = None
        if self.lap_method == "sign_inv":
            self.sign_inv_net = get_sign_inv_net(net_params)
        if self.pe_aggregate == "concat":
            self.pe_proj = nn.Linear(2 * hidden_dim, hidden_dim)

    def forward(self, g, h, p, e, snorm_n=None):
        if not (torch.is_tensor(h) and h.is_cuda):
            raise ValueError("GatedGCNNet inputs must be CUDA tensors (no CPU fallback)")
        gi = _graph_index(g, h.device)
        hd = self.hidden_dim
        x = EmbeddingSumFn.apply(h.to(torch.int64), self.embedding_h.weight)                 # [N, pad4(hidden)]
        pp = linear(p.reshape(p.shape[0], -1).contiguous(), self.embedding_p.weight, self.embedding_p.bias, pad4(hd))
        if self.pe_aggregate == "concat":                                                     # gatedgcn_net.py:98-101
            x = Linear2Fn.apply(x, pp, self.pe_proj.weight, self.pe_proj.bias, hd, hd)
        else:
            x = add_rows(x, pp)                                                               # h = h + p (:103)
        if self.edge_feat:
            ee = EmbeddingSumFn.apply(e.reshape(-1).to(torch.int64), self.embedding_e.weight)  # [E, pad4(hidden)]
        else:                                                                                 # e = Linear(1)(ones) (:106-108)
            ones = torch.ones(gi.E, 1, dtype=torch.float32, device=h.device)
            ee = linear(ones, self.embedding_e.weight, self.embedding_e.bias, pad4(hd))
        for layer in self.layers:
            x, ee = layer.forward_rows(gi, x, ee)
        if self.readout == "max":
            raise NotImplementedError("max readout is not built (no shipped configuration selects it)")
        hg = SegmentPoolFn.apply(x, gi, self.out_dim, self.readout != "sum")                  # mean (default) or sum
        self.g = g
        return self.MLP_layer(hg), g

    def loss(self, scores, targets):
        return torch.nn.functional.l1_loss(scores, targets)   # gatedgcn_net.py:156 (task loss)


def handle_lap(model, batch_pos_enc, batch_graphs, device=None, generator=None):
    """train/train_ZINC_graph_regression.py:12-47: the positional-encoding variants of the training loop.
    'sign_inv' runs the model's SignNet; 'sign_flip' / 'abs_val' / 'none' are element-wise; 'canonical' is sb_canonical_sign."""
    m = model.lap_method
    if m == "sign_flip":
        sign_flip = torch.rand(batch_pos_enc.size(1), generator=generator).to(batch_pos_enc.device)
        sign_flip = torch.where(sign_flip >= 0.5, 1.0, -1.0).to(batch_pos_enc.dtype)
        return batch_pos_enc * sign_flip.unsqueeze(0)
    if m == "abs_val":
        return batch_pos_enc.abs()
    if m == "sign_inv":
        return model.sign_inv_net(batch_graphs, batch_pos_enc.unsqueeze(-1)).squeeze(-1)
    if m == "canonical":
        if not batch_pos_enc.is_cuda:
            raise ValueError("handle_lap('canonical') expects a CUDA positional encoding (no CPU fallback)")
        pe = batch_pos_enc.contiguous().to(torch.float32)
        gi = _graph_index(batch_graphs, pe.device)
        out = torch.empty_like(pe)
        _call("sb_canonical_sign", _p(pe), pe.stride(0), _p(gi.graph_ptr), gi.B, pe.shape[1], _p(out), out.stride(0))
        return out
    if m == "none":
        return batch_pos_enc
    raise ValueError("invalid laplacian method")

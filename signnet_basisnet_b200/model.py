"""Downstream GNN predictor of the PyG trees (GINE stack + add-pool + output MLP) on the B200 kernels.

Mirrors Alchemy/sign_net/model.py:9-64, model_utils/pyg_gnn_wrapper.py:7-28 (GINConv, GINEConv) and
model_utils/elements.py:11-69 (Identity, DiscreteEncoder, MLP): same constructor arguments, same state_dict keys.
GAT/GCN/SimplifiedPNA wrappers are out of scope (never selected: main_alchemy.py:35 gnn_type='GINEConv').
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import counted_call as _call, ptr as _p
from .functional import batch_norm_act, linear, linear_fwd, linear_wgrad
from .layout import GraphIndex, pad4

BN = True


# ------------------------------------------------------------------------------------------------- autograd functions
class GineAggFn(torch.autograd.Function):
    """out_i = (1+eps) x_i + sum_{j->i} relu(x_j + e_ji) on padded [N, ld] rows (K6)."""

    @staticmethod
    def forward(ctx, x, e, eps, gi):
        x, e = x.contiguous(), e.contiguous()
        N, ld = x.shape
        out = torch.empty_like(x)
        _call("sb_gine_agg_fwd", _p(x), _p(e), _p(eps), _p(gi.in_ptr), _p(gi.in_src), _p(gi.in_eid), N, ld, _p(out))
        ctx.save_for_backward(x, e, eps)
        ctx.gi = gi
        return out

    @staticmethod
    def backward(ctx, gout):
        x, e, eps = ctx.saved_tensors
        gi = ctx.gi
        gout = gout.contiguous()
        N, ld = x.shape
        dx = torch.empty_like(x)
        de = torch.empty_like(e) if ctx.needs_input_grad[1] else None
        deps = torch.zeros(1, dtype=torch.float64, device=x.device)
        _call("sb_gine_agg_bwd", _p(gout), _p(x), _p(e), _p(eps), _p(gi.edge_index), _p(gi.out_ptr), _p(gi.out_dst),
              _p(gi.out_eid), N, gi.E, ld, _p(dx), _p(de), _p(deps))
        return dx, de, deps.to(torch.float32), None


class SegmentPoolFn(torch.autograd.Function):
    """scatter(x, batch, reduce='add'|'mean') over the sorted batch vector (K5)."""

    @staticmethod
    def forward(ctx, x, gi, C, mean):
        x = x.contiguous()
        out = torch.empty(gi.B, pad4(C), dtype=torch.float32, device=x.device)
        _call("sb_segment_pool_fwd", _p(x), x.stride(0), _p(gi.graph_ptr), gi.B, C, int(mean), _p(out), out.stride(0))
        ctx.cfg = (gi, C, mean, x.shape)
        return out

    @staticmethod
    def backward(ctx, gout):
        gi, C, mean, shape = ctx.cfg
        gout = gout.contiguous()
        gx = torch.empty(shape, dtype=torch.float32, device=gout.device)
        _call("sb_segment_pool_bwd", _p(gout), gout.stride(0), _p(gi.batch), _p(gi.graph_ptr), gi.N, C, int(mean),
              _p(gx), shape[1])
        return gx, None, None, None


class EmbeddingSumFn(torch.autograd.Function):
    """DiscreteEncoder: out[m] = sum_f table_f[idx[m, f]] (elements.py:31-37) -> padded [M, pad4(C)]."""

    @staticmethod
    def forward(ctx, idx, *tables):
        if idx.dim() == 1:
            idx = idx.unsqueeze(1)
        idx = idx.contiguous()
        if not (idx.is_cuda and idx.dtype == torch.int64):
            raise ValueError("DiscreteEncoder input must be a CUDA int64 tensor")
        M, F = idx.shape
        V, C = tables[0].shape
        out = torch.empty(M, pad4(C), dtype=torch.float32, device=idx.device)
        flags = torch.zeros(1, dtype=torch.int32, device=idx.device)
        for f in range(F):
            _call("sb_embedding_fwd", idx.data_ptr() + 8 * f, F, _p(tables[f]), V, C, M, _p(out), out.stride(0),
                  int(f > 0), _p(flags))
        ctx.idx, ctx.F, ctx.V, ctx.C, ctx.n_tables = idx, F, V, C, len(tables)
        ctx.flags = flags
        return out

    @staticmethod
    def backward(ctx, gout):
        gout = gout.contiguous()
        idx, F, V, C = ctx.idx, ctx.F, ctx.V, ctx.C
        ws = torch.empty(int(_lib.lib().sb_embedding_bwd_workspace_floats(V, C)), dtype=torch.float32,
                         device=gout.device)
        grads = []
        for f in range(ctx.n_tables):
            if f < F:
                g = torch.empty(V, C, dtype=torch.float32, device=gout.device)
                _call("sb_embedding_bwd", idx.data_ptr() + 8 * f, F, _p(gout), gout.stride(0), V, C, idx.shape[0],
                      _p(g), _p(ws))
                grads.append(g)
            else:
                grads.append(None)
        return (None, *grads)


class Linear2Fn(torch.autograd.Function):
    """y = [x1, x2] @ W^T + b without materialising the concatenation (model.py:40: Linear(cat[x, pos]))."""

    @staticmethod
    def forward(ctx, x1, x2, W, b, K1, K2):
        x1, x2, W = x1.contiguous(), x2.contiguous(), W.contiguous()
        M, N, K = x1.shape[0], W.shape[0], W.shape[1]
        y = torch.empty(M, pad4(N), dtype=torch.float32, device=x1.device)
        linear_fwd(x1, x1.stride(0), W, K, 1, None, y, y.stride(0), M, 1, K1, N)
        linear_fwd(x2, x2.stride(0), W[:, K1:], K, 1, b, y, y.stride(0), M, 1, K2, N, accumulate=True)
        ctx.save_for_backward(x1, x2, W)
        ctx.dims = (K1, K2, b is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x1, x2, W = ctx.saved_tensors
        K1, K2, has_b = ctx.dims
        gy = gy.contiguous()
        M, N, K = x1.shape[0], W.shape[0], W.shape[1]
        g1 = torch.empty_like(x1)
        g2 = torch.empty_like(x2)
        linear_fwd(gy, gy.stride(0), W, 1, K, None, g1, x1.shape[1], M, 1, N, K1)
        linear_fwd(gy, gy.stride(0), W[:, K1:], 1, K, None, g2, x2.shape[1], M, 1, N, K2)
        gW = torch.empty_like(W)
        gb = torch.empty(N, dtype=torch.float32, device=gy.device) if has_b else None
        linear_wgrad(gy, gy.stride(0), x1, x1.stride(0), M, 1, N, K1, gW, K, 1, gb)
        linear_wgrad(gy, gy.stride(0), x2, x2.stride(0), M, 1, N, K2, gW[:, K1:], K, 1, None)
        return g1, g2, gW, gb, None, None


# ------------------------------------------------------------------------------------------------------------ modules
class Identity(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, input):
        return input

    def reset_parameters(self):
        pass


class DiscreteEncoder(nn.Module):
    def __init__(self, hidden_channels, max_num_features=10, max_num_values=6):
        super().__init__()
        self.embeddings = nn.ModuleList([nn.Embedding(max_num_values, hidden_channels)
                                         for _ in range(max_num_features)])

    def reset_parameters(self):
        for embedding in self.embeddings:
            embedding.reset_parameters()

    def forward(self, x):
        return EmbeddingSumFn.apply(x, *[emb.weight for emb in self.embeddings])


class MLP(nn.Module):
    """elements.MLP (:39-69): Linear (+BN +ReLU) x nlayer on [M, d] rows; returns the padded [M, pad4(nout)]."""

    def __init__(self, nin, nout, nlayer=2, with_final_activation=True, with_norm=BN, bias=True, nhid=None):
        super().__init__()
        n_hid = nin if nhid is None else nhid
        self.layers = nn.ModuleList([
            nn.Linear(nin if i == 0 else n_hid, n_hid if i < nlayer - 1 else nout,
                      bias=True if (i == nlayer - 1 and not with_final_activation and bias) or (not with_norm) else False)
            for i in range(nlayer)])
        self.norms = nn.ModuleList([nn.BatchNorm1d(n_hid if i < nlayer - 1 else nout) if with_norm else Identity()
                                    for i in range(nlayer)])
        self.nlayer = nlayer
        self.with_final_activation = with_final_activation
        self.with_norm = with_norm
        self.residual = (nin == nout)

    def reset_parameters(self):
        for layer, norm in zip(self.layers, self.norms):
            layer.reset_parameters()
            norm.reset_parameters()

    def forward(self, x):
        for i, (layer, norm) in enumerate(zip(self.layers, self.norms)):
            x = linear(x, layer.weight, layer.bias, pad4(layer.out_features))
            if i < self.nlayer - 1 or self.with_final_activation:
                if self.with_norm:
                    x = batch_norm_act(x, norm, self.training, relu=True)
                else:
                    raise NotImplementedError("MLP(with_norm=False) with activation is not used on the hot path")
        return x


class _Eps(nn.Module):
    def __init__(self):
        super().__init__()
        self.eps = nn.Parameter(torch.Tensor([0.0]))
        self.nn = None

    def reset_parameters(self):
        self.eps.data.fill_(0.0)


class GINEConv(nn.Module):
    """pyg_gnn_wrapper.GINEConv (:19-28): nn((1+eps) x_i + sum relu(x_j + e_ij)).  The PyG layer registers the MLP
    both as `nn` and as `layer.nn`; both key sets appear in the reference state_dict, so both exist here."""

    def __init__(self, nin, nout, bias=True):
        super().__init__()
        self.nn = MLP(nin, nout, 2, False, bias=bias)
        self.layer = _Eps()
        self.layer.nn = self.nn

    def reset_parameters(self):
        self.nn.reset_parameters()
        self.layer.reset_parameters()

    def forward(self, x, gi: GraphIndex, edge_emb):
        return self.nn(GineAggFn.apply(x, edge_emb, self.layer.eps, gi))


class GNN(nn.Module):
    def __init__(self, nfeat_node, nfeat_edge, nhid, nout, nlayer, gnn_type="GINEConv", dropout=0, pooling="add",
                 bn=BN, res=True, max_num_values=6):
        super().__init__()
        if gnn_type != "GINEConv":
            raise ValueError("only GINEConv is on the SignNet hot path (main_alchemy.py:35)")
        if pooling != "add":
            raise NotImplementedError("pooling='add' is the only read-out the shipped configurations use")
        if dropout != 0:
            raise NotImplementedError("dropout is 0 in every shipped configuration")
        mk_disc = lambda: DiscreteEncoder(nhid, max_num_values=max_num_values)
        self.input_encoder = mk_disc() if nfeat_node is None else MLP(nfeat_node, nhid, 1)
        self.edge_encoders = nn.ModuleList([mk_disc() if nfeat_edge is None else MLP(nfeat_edge, nhid, 1)
                                            for _ in range(nlayer)])
        self.convs = nn.ModuleList([GINEConv(nhid, nhid, bias=not bn) for _ in range(nlayer)])
        self.norms = nn.ModuleList([nn.BatchNorm1d(nhid) if bn else Identity() for _ in range(nlayer)])
        self.output_encoder = MLP(nhid, nout, nlayer=2, with_final_activation=False,
                                  with_norm=False if pooling == "mean" else True)
        if max_num_values != 6:  # GINESignNetPyG flavour: allocated, used only for mean pooling (core/model.py:18,72)
            self.size_embedder = nn.Embedding(200, nhid)
        self.linear = nn.Linear(2 * nhid, nhid)
        self.pooling, self.dropout, self.res, self.bn = pooling, dropout, res, bn
        self.nhid, self.nout = nhid, nout

    def reset_parameters(self):
        self.input_encoder.reset_parameters()
        self.output_encoder.reset_parameters()
        self.linear.reset_parameters()
        for edge_encoder, conv, norm in zip(self.edge_encoders, self.convs, self.norms):
            edge_encoder.reset_parameters()
            conv.reset_parameters()
            norm.reset_parameters()

    def forward(self, data, additional_x=None, graph_index=None):
        gi = graph_index or getattr(data, "_b200_graph_index", None) or GraphIndex(
            data.edge_index, data.batch, getattr(data, "num_graphs", None))
        return self.forward_tensors(data.x, data.edge_index, data.edge_attr, data.batch, additional_x, gi)

    def forward_tensors(self, x_in, edge_index, edge_attr, batch, additional_x, gi: GraphIndex):
        """GNN.forward (model.py:36-64)."""
        d = self.nhid
        x_in = x_in.squeeze() if x_in.dim() > 1 and x_in.shape[-1] == 1 else x_in
        x = self.input_encoder(x_in if x_in.is_floating_point() else x_in.contiguous())
        if additional_x is not None:
            x = Linear2Fn.apply(x, additional_x, self.linear.weight, self.linear.bias, d, d)
        if edge_attr is None:
            edge_attr = edge_index.new_zeros(edge_index.size(-1))
        prev = x
        for edge_encoder, conv, norm in zip(self.edge_encoders, self.convs, self.norms):
            e = edge_encoder(edge_attr)
            x = conv(x, gi, e)
            if self.bn:
                x = batch_norm_act(x, norm, self.training, relu=True, res=prev if self.res else None)
            else:
                raise NotImplementedError("GNN(bn=False) is not used on the hot path")
            prev = x
        x = SegmentPoolFn.apply(x, gi, d, self.pooling == "mean")
        x = self.output_encoder(x)
        return x[:, :self.nout] if x.shape[1] != self.nout else x

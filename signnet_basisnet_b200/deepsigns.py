"""DGL-flavour SignNets (`sign_inv_net` in {gin, masked_gin}) on the B200 kernels.

Mirrors GraphPrediction/layers/deepsigns.py:33-86 (GINDeepSigns, MaskedGINDeepSigns), layers/gnns.py:81-114 (GIN),
layers/mlp.py:5-56 (MLP) and nets/ZINC_graph_regression/sign_inv_net.py:3-18 (get_sign_inv_net) with the reference's
constructor arguments and state_dict keys (`enc.layers.{l}.apply_func.lins.*`, `enc.layers.{l}.eps` buffer,
`enc.bns.*`, `rho.lins.*`, `rho.bns.*`).  DGL itself is absent: `g` is any object exposing `edges() -> (src, dst)` and
`batch_num_nodes()` (a DGLGraph does).  In this flavour every one of the k slots is live (BatchNorm statistics include
the zero-padded columns, gnns.py:107-112), so the slot-row layout is built unmasked; both sign passes run side by side.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import counted_call as _call, ptr as _p
from .functional import batch_norm_act, linear, slot_sum
from .layout import GraphIndex, pad4
from .phi import gin_agg


class RowsAggFn(torch.autograd.Function):
    """(1+eps) x + sum_nbr x on [S, R, ld] slot rows (dgl GINConv 'sum', eps buffer)."""

    @staticmethod
    def forward(ctx, x, eps, slots):
        x = x.contiguous()
        S = x.shape[0]
        ld = 1 if x.dim() == 2 else x.shape[2]
        out = torch.empty_like(x)
        gin_agg(x, out, slots, S, ld, eps=eps)
        ctx.cfg = (slots, S, ld, eps)
        return out

    @staticmethod
    def backward(ctx, g):
        slots, S, ld, eps = ctx.cfg
        g = g.contiguous()
        gx = torch.empty_like(g)
        gin_agg(g, gx, slots, S, ld, eps=eps, transpose=True)
        return gx, None, None


class DenseToRowsFn(torch.autograd.Function):
    """x [N, k, C] -> rows [2, R, ld] holding (+x, -x) (the two sign passes, deepsigns.py:46,73)."""

    @staticmethod
    def forward(ctx, x, slots):
        gi = slots.gi
        x = x.contiguous()
        C = x.shape[2]
        ld = pad4(C) if C > 1 else 1
        rows = torch.empty((2, slots.R) if ld == 1 else (2, slots.R, ld), dtype=torch.float32, device=x.device)
        _call("sb_dense_to_rows", _p(x), slots.R, 2, 1, _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr), gi.N, slots.k,
              int(slots.masked), C, ld, _p(rows))
        return rows

    @staticmethod
    def backward(ctx, g):
        return None, None   # eigenvector inputs are data, never differentiated


class RowsToDenseFn(torch.autograd.Function):
    """rows [S, R, ld] -> [N, k*C] summed over the S sign passes (enc(x) + enc(-x) then reshape, deepsigns.py:46-48)."""

    @staticmethod
    def forward(ctx, rows, slots, C):
        gi = slots.gi
        rows = rows.contiguous()
        S, R, ld = rows.shape
        out = torch.empty(gi.N, slots.k * C, dtype=torch.float32, device=rows.device)
        _call("sb_rows_to_dense", _p(rows), ld, R, S, _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr), gi.N, slots.k,
              int(slots.masked), C, _p(out))
        ctx.cfg = (slots, S, R, ld, C)
        return out

    @staticmethod
    def backward(ctx, g):
        slots, S, R, ld, C = ctx.cfg
        gi = slots.gi
        g = g.contiguous()
        rows = torch.zeros(S, R, ld, dtype=torch.float32, device=g.device)
        _call("sb_dense_to_rows", _p(g), R, S, 0, _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr), gi.N, slots.k,
              int(slots.masked), C, ld, _p(rows))
        return rows, None, None


class MLP(nn.Module):
    """layers/mlp.py MLP: (Linear -> activation -> BN -> dropout) x (L-1) -> Linear."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, use_bn=False, use_ln=False, dropout=0.5,
                 activation="relu", residual=False):
        super().__init__()
        if use_ln or residual or activation != "relu":
            raise NotImplementedError("shipped sign_inv configurations use relu + BatchNorm, no LayerNorm/residual")
        if dropout != 0:
            raise NotImplementedError("dropout is 0.0 in every shipped sign_inv configuration")
        self.lins = nn.ModuleList()
        if use_bn:
            self.bns = nn.ModuleList()
        if num_layers == 1:
            self.lins.append(nn.Linear(in_channels, out_channels))
        else:
            self.lins.append(nn.Linear(in_channels, hidden_channels))
            if use_bn:
                self.bns.append(nn.BatchNorm1d(hidden_channels))
            for _ in range(num_layers - 2):
                self.lins.append(nn.Linear(hidden_channels, hidden_channels))
                if use_bn:
                    self.bns.append(nn.BatchNorm1d(hidden_channels))
            self.lins.append(nn.Linear(hidden_channels, out_channels))
        self.use_bn, self.dropout, self.residual = use_bn, dropout, residual

    def forward(self, x, G=1):
        """x [M, ld] (G = 1) or [G, M, ld] -> same leading shape, padded output columns."""
        lead = x.shape[:-1]
        for i, lin in enumerate(self.lins[:-1]):
            y = linear(x.reshape(-1, x.shape[-1]), lin.weight, lin.bias, pad4(lin.out_features), relu=True)
            x = y.reshape(*lead, y.shape[-1])
            if self.use_bn:
                x = batch_norm_act(x, self.bns[i], self.training, relu=False, G=G)
        lin = self.lins[-1]
        y = linear(x.reshape(-1, x.shape[-1]), lin.weight, lin.bias, pad4(lin.out_features))
        return y.reshape(*lead, y.shape[-1])


class GINLayer(nn.Module):
    """Parameter holder of dgl.nn.pytorch.GINConv(apply_func, 'sum'): `apply_func` + a fixed-zero `eps` buffer."""

    def __init__(self, apply_func):
        super().__init__()
        self.apply_func = apply_func
        self.register_buffer("eps", torch.FloatTensor([0]))


class GIN(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels, n_layers, use_bn=True, dropout=0.5,
                 activation="relu"):
        super().__init__()
        if dropout != 0 or activation != "relu":
            raise NotImplementedError("shipped sign_inv configurations use relu and dropout 0.0")
        self.layers = nn.ModuleList()
        if use_bn:
            self.bns = nn.ModuleList()
        self.use_bn = use_bn
        mk = lambda i, o: GINLayer(MLP(i, hidden_channels, o, 2, use_bn=use_bn, dropout=dropout, activation=activation))
        self.layers.append(mk(in_channels, hidden_channels))
        for _ in range(n_layers - 2):
            self.layers.append(mk(hidden_channels, hidden_channels))
            if use_bn:
                self.bns.append(nn.BatchNorm1d(hidden_channels))
        self.layers.append(mk(hidden_channels, out_channels))
        if use_bn:
            self.bns.append(nn.BatchNorm1d(hidden_channels))

    def forward_rows(self, x_rows, slots):
        """x_rows [2, R] (+x, -x) -> [2, R, pad4(out)]  (GIN.forward gnns.py:102-114, both signs at once)."""
        x = x_rows
        for i, layer in enumerate(self.layers):
            if i != 0 and self.use_bn:
                x = batch_norm_act(x, self.bns[i - 1], self.training, relu=False, G=2)
            x = RowsAggFn.apply(x, layer.eps, slots)
            if x.dim() == 2:
                x = x.unsqueeze(-1)
            x = layer.apply_func(x, G=2)
        return x


def _graph_index(g, device):
    src, dst = g.edges()
    n = torch.as_tensor(g.batch_num_nodes()).to(device=device, dtype=torch.int64)
    batch = torch.repeat_interleave(torch.arange(n.numel(), device=device), n)
    edge_index = torch.stack([src.to(device=device, dtype=torch.int64), dst.to(device=device, dtype=torch.int64)])
    return GraphIndex(edge_index, batch, int(n.numel()))


class GINDeepSigns(nn.Module):
    """f(v1..vk) = rho([enc(vi) + enc(-vi)]_i), rho = MLP on the concatenation (deepsigns.py:33-51)."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, k, use_bn=False, use_ln=False,
                 dropout=0.5, activation="relu"):
        super().__init__()
        self.enc = GIN(in_channels, hidden_channels, out_channels, num_layers, use_bn=use_bn, dropout=dropout,
                       activation=activation)
        self.rho = MLP(out_channels * k, hidden_channels, k, num_layers, use_bn=use_bn, dropout=dropout,
                       activation=activation)
        self.k, self.out_channels = k, out_channels

    def forward(self, g, x):
        gi = _graph_index(g, x.device)
        slots = gi.slots(self.k, False, pad4(self.out_channels))
        rows = DenseToRowsFn.apply(x, slots)
        h = self.enc.forward_rows(rows, slots)
        h = RowsToDenseFn.apply(h, slots, self.out_channels)
        y = self.rho(h)
        return y[:, :self.k].reshape(x.shape[0], self.k, 1)


class MaskedGINDeepSigns(nn.Module):
    """rho(sum_{i < n_b} enc(vi) + enc(-vi)): the north-star formula (deepsigns.py:54-86)."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, k, device=None, use_bn=False,
                 use_ln=False, dropout=0.5, activation="relu"):
        super().__init__()
        self.device = device
        self.enc = GIN(in_channels, hidden_channels, out_channels, num_layers, use_bn=use_bn, dropout=dropout,
                       activation=activation)
        self.rho = MLP(out_channels, hidden_channels, k, num_layers, use_bn=use_bn, dropout=dropout,
                       activation=activation)
        self.k, self.out_channels = k, out_channels

    def forward(self, g, x):
        gi = _graph_index(g, x.device)
        slots = gi.slots(self.k, False, pad4(self.out_channels))
        rows = DenseToRowsFn.apply(x, slots)
        h = self.enc.forward_rows(rows, slots)
        h = slot_sum(h, slots, self.out_channels, limit_by_n=True)   # x[~mask] = 0 ; x.sum(dim=1)
        y = self.rho(h)
        return y[:, :self.k].reshape(x.shape[0], self.k, 1)


def get_sign_inv_net(net_params):
    """nets/ZINC_graph_regression/sign_inv_net.py:3-18 (gcn / transformer variants are out of scope: no shipped
    configuration selects them)."""
    assert net_params["sign_inv_net"] is not None, "did not specify sign inv net"
    kind = net_params["sign_inv_net"]
    args = (1, net_params["hidden_dim"], net_params["phi_out_dim"], net_params["sign_inv_layers"],
            net_params["pos_enc_dim"])
    kw = dict(use_bn=True, dropout=net_params["dropout"], activation=net_params["sign_inv_activation"])
    if kind == "gin":
        return GINDeepSigns(*args, **kw)
    if kind == "masked_gin":
        return MaskedGINDeepSigns(*args, net_params["device"], **kw)
    raise ValueError("Invalid sign inv net")

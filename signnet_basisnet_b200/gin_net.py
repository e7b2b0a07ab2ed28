"""DGL-flavour GIN predictor consuming the sign-invariant positional encoding.

Mirrors GraphPrediction/nets/ZINC_graph_regression/gin_net.py:19-138 (`GINNet`, the `pe_init='lap_pe'` / no-LSPE path
every shipped `*_signinv_*` GIN configuration uses) and layers/mlp_readout_layer.py:9-25 (`MLPReadout`): same
`net_params` keys, same state_dict keys (`embedding_h`, `embedding_p`, `embedding_e`, `layers.{l}.apply_func.*`,
`layers.{l}.eps` buffer, `MLP_layer.FC_layers.*`, `sign_inv_net.*`), `forward(g, h, p, e, snorm_n) -> (scores, g)` as
called at train/train_ZINC_graph_regression.py:76.  The same kernels as phi with a single slot per node: dgl
GINConv('sum') = sb_gin_agg on k = 1 slot rows (rows == nodes), MLP = tcgen05 / FFMA Linear + BatchNorm kernels,
readout = sb_segment_pool.  The LSPE branch (`pe_init='rand_walk'` / `lap_lspe`) is out of scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .deepsigns import GINLayer, MLP, RowsAggFn, _graph_index, get_sign_inv_net
from .functional import add_rows, linear
from .layout import pad4
from .model import EmbeddingSumFn, SegmentPoolFn


class MLPReadout(nn.Module):
    """layers/mlp_readout_layer.py:9-25: L halving Linear+ReLU layers, then Linear to output_dim."""

    def __init__(self, input_dim, output_dim, L=2):
        super().__init__()
        fcs = [nn.Linear(input_dim // 2 ** l, input_dim // 2 ** (l + 1), bias=True) for l in range(L)]
        fcs.append(nn.Linear(input_dim // 2 ** L, output_dim, bias=True))
        self.FC_layers = nn.ModuleList(fcs)
        self.L = L

    def forward(self, x):
        y = x
        for l in range(self.L):
            fc = self.FC_layers[l]
            y = linear(y, fc.weight, fc.bias, pad4(fc.out_features), relu=True)
        fc = self.FC_layers[self.L]
        return linear(y, fc.weight, fc.bias, pad4(fc.out_features))[:, :fc.out_features]


class GINNet(nn.Module):
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
        self.pos_enc_dim = net_params["pos_enc_dim"]
        if self.pe_init == "rand_walk" or self.lap_lspe:
            raise NotImplementedError("GINNet on the B200 path: the LSPE branch is out of scope (SURVEY §8 row a13)")
        if net_params["in_feat_dropout"] != 0 or net_params["dropout"] != 0:
            raise NotImplementedError("dropout is 0.0 in every shipped sign_inv configuration")
        if self.pe_init in ("rand_walk", "lap_pe"):
            self.embedding_p = nn.Linear(self.pos_enc_dim, hidden_dim)
        self.embedding_h = nn.Embedding(net_params["num_atom_type"], hidden_dim)
        # allocated by the reference, evaluated and discarded in its forward (dgl GINConv takes no edge feature)
        self.embedding_e = (nn.Embedding(net_params["num_bond_type"], hidden_dim) if self.edge_feat
                            else nn.Linear(1, hidden_dim))
        mk = lambda o: GINLayer(MLP(hidden_dim, hidden_dim, o, 2, use_bn=self.batch_norm, dropout=0.0, activation="relu"))
        self.layers = nn.ModuleList([mk(hidden_dim) for _ in range(self.n_layers - 1)] + [mk(out_dim)])
        self.MLP_layer = MLPReadout(out_dim, 1)
        self.out_dim = out_dim
        self.g = None
        if self.lap_method == "sign_inv":
            self.sign_inv_net = get_sign_inv_net(net_params)

    def forward(self, g, h, p, e, snorm_n=None):
        if not (torch.is_tensor(h) and h.is_cuda):
            raise ValueError("GINNet inputs must be CUDA tensors (no CPU fallback)")
        gi = _graph_index(g, h.device)
        hd = self.embedding_h.embedding_dim
        x = EmbeddingSumFn.apply(h.to(torch.int64), self.embedding_h.weight)          # [N, pad4(hidden)]
        if self.pe_init in ("rand_walk", "lap_pe"):
            pp = linear(p.reshape(p.shape[0], -1).contiguous(), self.embedding_p.weight, self.embedding_p.bias, pad4(hd))
            x = add_rows(x, pp)                                                        # h = h + p (gin_net.py:90-92)
        slots = gi.slots(1, False, x.shape[1])                                         # one slot per node: rows == nodes
        x = x.unsqueeze(0)
        for layer in self.layers:
            x = RowsAggFn.apply(x, layer.eps, slots)                                   # dgl GINConv(.., 'sum')
            x = layer.apply_func(x, G=1)
        x = x.squeeze(0)
        hg = SegmentPoolFn.apply(x, gi, self.out_dim, self.readout != "sum")           # mean (default) or sum readout
        if self.readout == "max":
            raise NotImplementedError("max readout is not built (no shipped configuration selects it)")
        self.g = g
        return self.MLP_layer(hg), g

    def loss(self, scores, targets):
        return torch.nn.functional.l1_loss(scores, targets)   # gin_net.py:143 (task loss; lapeig loss out of scope)

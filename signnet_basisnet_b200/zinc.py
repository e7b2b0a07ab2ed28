"""The GINESignNetPyG tree's module surface under its own constructor signatures
(GINESignNetPyG/core/sign_net.py:123 `SignNetGNN(node_feat, edge_feat, n_hid, n_out, nl_signnet, nl_gnn)`, :80
`SignNet(n_hid, nl_phi, nl_rho=2)`, :12 `GNN3d(n_in, n_out, n_layer, gnn_type)`): a training script of that tree
(`train/zinc.py:60,76`) switches with `from signnet_basisnet_b200.zinc import SignNetGNN` and no other change.

Same classes as signnet_basisnet_b200.sign_net with the tree's parameterisation fixed: phi MLP hidden width = n_in and no
bias on its second Linear (core/model_utils/masked_layers.py:66-69), nl_rho = 1 (core/sign_net.py:125), eigenvalue
encoders allocated but unused (quirk v), DiscreteEncoder inputs when node_feat / edge_feat are None."""
from __future__ import annotations

from . import sign_net as _sn


class GNN3d(_sn.GNN3d):
    def __init__(self, n_in, n_out, n_layer, gnn_type="MaskedGINConv"):
        super().__init__(n_in, n_out, n_layer, gnn_type=gnn_type, flavour="zinc")


class SignNet(_sn.SignNet):
    def __init__(self, n_hid, nl_phi, nl_rho=2):
        super().__init__(n_hid, nl_phi, nl_rho=nl_rho, flavour="zinc")


class SignNetGNN(_sn.SignNetGNN):
    def __init__(self, node_feat, edge_feat, n_hid, n_out, nl_signnet, nl_gnn):
        super().__init__(node_feat, edge_feat, n_hid, n_out, nl_signnet, nl_gnn, flavour="zinc")

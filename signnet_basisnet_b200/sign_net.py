"""PyG-flavour SignNet modules with the reference's constructor signatures and state_dict keys, executing on the
hand-written sm_100a kernels of libsignnet_b200 (no torch_geometric / torch_scatter, no CPU path).

Mirrors Alchemy/sign_net/sign_net.py:12-132 and model_utils/masked_layers.py:7-84; `flavour='zinc'` selects the
GINESignNetPyG parameterisation (core/sign_net.py:18-22,123-126; core/model_utils/masked_layers.py:66-69): phi MLP
hidden = n_in, no bias on the second Linear, nl_rho = 1.

Drop-in surface:
    SignNetGNN(node_feat, edge_feat, n_hid, n_out, nl_signnet, nl_gnn, nl_rho=4, ignore_eigval=False,
               gnn_type='GINEConv').forward(data)            -> [B, n_out]
    SignNet(n_hid, nl_phi, nl_rho=2, ignore_eigval=False).forward(data)          -> [N, n_hid]
plus the tensor-level overload forward(x, edge_index, eigvecs[N,k], batch, edge_attr=None, eigvals=None).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import counted_call as _call, ptr as _p
from .functional import batch_norm_act, linear, slot_sum
from .layout import GraphIndex, pad4
from .phi import PhiStackFn


class Identity(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, input):
        return input

    def reset_parameters(self):
        pass


class MaskedBN(nn.Module):
    """Parameter holder for masked_layers.MaskedBN (:7-20); on slot rows every row is valid, so it is plain BN."""

    def __init__(self, num_features):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features)

    def reset_parameters(self):
        self.bn.reset_parameters()

    def forward(self, x_rows, relu=False):
        return batch_norm_act(x_rows, self.bn, self.training, relu=relu)


class MaskedMLP(nn.Module):
    """masked_layers.MaskedMLP (:34-64): same layers/norms (including the final norm that is allocated but never
    applied when with_final_activation=False, reference quirk ii)."""

    def __init__(self, nin, nout, nlayer=2, with_final_activation=True, with_norm=True, bias=True, nhid=None):
        super().__init__()
        n_hid = nin if nhid is None else nhid
        self.layers = nn.ModuleList([
            nn.Linear(nin if i == 0 else n_hid, n_hid if i < nlayer - 1 else nout,
                      bias=True if (i == nlayer - 1 and not with_final_activation and bias) or (not with_norm) else False)
            for i in range(nlayer)])
        self.norms = nn.ModuleList([MaskedBN(n_hid if i < nlayer - 1 else nout) if with_norm else Identity()
                                    for i in range(nlayer)])
        self.nlayer = nlayer
        self.with_final_activation = with_final_activation
        self.with_norm = with_norm
        self.residual = (nin == nout)

    def reset_parameters(self):
        for layer, norm in zip(self.layers, self.norms):
            layer.reset_parameters()
            norm.reset_parameters()

    def forward(self, x_rows):
        """x_rows [M, >= nin] (all rows valid) -> [M, pad4(nout)]."""
        x = x_rows
        for i, (layer, norm) in enumerate(zip(self.layers, self.norms)):
            x = linear(x, layer.weight, layer.bias, pad4(layer.out_features))
            if i < self.nlayer - 1 or self.with_final_activation:
                if self.with_norm:
                    x = norm(x, relu=True)
                else:
                    raise NotImplementedError("MaskedMLP without norm is never instantiated by the reference")
        return x


class GINEps(nn.Module):
    """Holds the learnable eps of gnn.GINConv(Identity(), train_eps=True) under the reference key `layer.eps`."""

    def __init__(self, eps=0.0, train_eps=True):
        super().__init__()
        self.initial_eps = eps
        if train_eps:
            self.eps = nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))

    def reset_parameters(self):
        self.eps.data.fill_(self.initial_eps)


class MaskedGINConv(nn.Module):
    """masked_layers.MaskedGINConv (:66-84): parameter holder; executed inside PhiStackFn."""

    def __init__(self, nin, nout, bias=True, nhid=None):
        super().__init__()
        self.nn = MaskedMLP(nin, nout, 2, False, bias=bias, nhid=nhid)
        self.layer = GINEps(train_eps=True)

    def reset_parameters(self):
        self.nn.reset_parameters()
        self.layer.reset_parameters()


class GNN3d(nn.Module):
    """sign_net.GNN3d (:12-44): L x {MaskedGINConv -> MaskedBN -> ReLU -> residual} on k independent slot channels."""

    def __init__(self, n_in, n_out, n_layer, gnn_type="MaskedGINConv", flavour="alchemy"):
        super().__init__()
        if gnn_type != "MaskedGINConv":
            raise ValueError("only MaskedGINConv is on the SignNet hot path")
        if flavour == "alchemy":   # Alchemy/sign_net/sign_net.py:20
            mk = lambda i: MaskedGINConv(n_in if i == 0 else n_out, n_out, bias=True, nhid=n_out)
        elif flavour == "zinc":    # GINESignNetPyG/core/sign_net.py:20
            mk = lambda i: MaskedGINConv(n_in if i == 0 else n_out, n_out, bias=False)
        else:
            raise ValueError(flavour)
        self.convs = nn.ModuleList([mk(i) for i in range(n_layer)])
        self.norms = nn.ModuleList([MaskedBN(n_out) for _ in range(n_layer)])
        if flavour == "zinc":  # allocated by the reference, never run (core/sign_net.py:22,37-39; quirk v)
            from .model import DiscreteEncoder
            self.edge_encoders = nn.ModuleList([DiscreteEncoder(n_in if i == 0 else n_out, max_num_values=500)
                                                for i in range(n_layer)])

    def reset_parameters(self):
        for conv, norm in zip(self.convs, self.norms):
            conv.reset_parameters()
            norm.reset_parameters()
        for enc in getattr(self, "edge_encoders", []):
            enc.reset_parameters()

    # ------------------------------------------------------------------------------------------------ execution
    def _flat_params(self):
        dims, params, buffers = [], [], []
        for conv, norm in zip(self.convs, self.norms):
            l0, l1 = conv.nn.layers
            bn0 = conv.nn.norms[0].bn
            dims.append((l0.in_features, l0.out_features, l1.out_features))
            params += [l0.weight, bn0.weight, bn0.bias, l1.weight, l1.bias, conv.layer.eps, norm.bn.weight,
                       norm.bn.bias]
            buffers.append((bn0.running_mean, bn0.running_var, norm.bn.running_mean, norm.bn.running_var))
        return dims, params, buffers

    def forward_rows(self, x0, gi: GraphIndex, k: int, masked: bool = True, capture=None):
        """x0 [2, R] (+v, -v in slot-row order) -> X_L [2, R, pad4(n_out)]."""
        dims, params, buffers = self._flat_params()
        d = dims[-1][2]
        slots = gi.slots(k, masked, pad4(d))
        slots_in = gi.slots(k, masked, dims[0][0])
        if self.training:  # one BatchNorm call per sign pass: two (+v, -v) under SignNet.forward (sign_net.py:113)
            counters = [b.num_batches_tracked for conv, norm in zip(self.convs, self.norms)
                        for b in (conv.nn.norms[0].bn, norm.bn) if b.num_batches_tracked is not None]
            if counters:   # one multi-tensor launch instead of one tiny kernel per BatchNorm
                torch._foreach_add_(counters, int(x0.shape[0]))
        cfg = dict(slots=slots, slots_in=slots_in, dims=dims, training=self.training, buffers=buffers,
                   capture=capture)
        return PhiStackFn.apply(x0, cfg, *params), slots

    def forward(self, x, edge_index, edge_attr=None, mask=None, batch=None, num_graphs=None):
        """Reference signature GNN3d.forward(x[N,k,n_in], edge_index, edge_attr, mask[N,k]) (sign_net.py:28-44): ONE
        sign pass, returns the dense [N, k, n_out] tensor with zeros in the masked slots.  `batch` (graph id per node,
        sorted) is additionally required: the reference reads graph membership implicitly from `mask`, the slot-row
        layout needs it explicitly.  `mask`, when given, must be the reference's own mask (slot j valid iff j < n_b,
        sign_net.py:100-102).  SignNet.forward does not come through here (it runs both sign passes side by side)."""
        if batch is None:
            raise ValueError("GNN3d.forward needs `batch` (graph id per node) on the B200 path")
        if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3):
            raise ValueError("x must be a CUDA float32 tensor [N, k, n_in] (no CPU fallback on the SignNet hot path)")
        from .deepsigns import RowsToDenseFn

        gi = GraphIndex(edge_index, batch, num_graphs)
        N, k, C = x.shape
        dims = self._flat_params()[0]
        if C != dims[0][0] or N != gi.N:
            raise ValueError(f"x is {tuple(x.shape)}; expected [{gi.N}, k, {dims[0][0]}]")
        d = dims[-1][2]
        masked = mask is not None
        slots = gi.slots(k, masked, pad4(d))
        if masked:
            n = (gi.graph_ptr[1:] - gi.graph_ptr[:-1]).long()
            want = torch.arange(k, device=x.device)[None, :] < n[gi.batch][:, None]
            if mask.shape != want.shape or not torch.equal(mask.bool(), want):
                raise ValueError("mask must be the SignNet slot mask: mask[i, j] = j < n_{batch[i]}")
        ld = pad4(C) if C > 1 else 1
        rows = torch.empty((1, slots.R) if ld == 1 else (1, slots.R, ld), dtype=torch.float32, device=x.device)
        _call("sb_dense_to_rows", _p(x.contiguous()), slots.R, 1, 0, _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr),
              gi.N, slots.k, int(slots.masked), C, ld, _p(rows))
        xr, slots = self.forward_rows(rows, gi, k, masked)
        return RowsToDenseFn.apply(xr, slots, d).view(N, k, d)


def build_phi_input(gi: GraphIndex, slots, eigen_vectors=None, eigvecs_dense=None):
    """x0 [2, R]: +v / -v in slot-row order, from the ragged PyG eigen data or a dense-list [N, >=k] tensor."""
    x0 = torch.empty(2, slots.R, dtype=torch.float32, device=gi.device)
    if eigvecs_dense is not None:
        ev = eigvecs_dense.contiguous()
        if ev.dim() != 2 or ev.shape[0] != gi.N or ev.shape[1] < slots.k:
            raise ValueError(f"eigvecs must be [N, >=k]; got {tuple(ev.shape)} for N={gi.N}, k={slots.k}")
        _call("sb_phi_input_dense", _p(ev), ev.shape[1], _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr), gi.N,
              slots.k, int(slots.masked), slots.R, _p(x0))
    else:
        ev = eigen_vectors.contiguous()
        if ev.numel() != slots.vec_total:
            raise ValueError(f"eigen_vectors has {ev.numel()} entries, batch needs sum n_b^2 = {slots.vec_total}")
        _call("sb_phi_input_ragged", _p(ev), _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr), _p(slots.vec_ptr),
              gi.N, slots.k, int(slots.masked), slots.R, _p(x0))
    return x0


class SetTransformer(nn.Module):
    """rho of the PyG trees (sign_net.py:46-72): nl_rho x TransformerEncoderLayer over the slot tokens of every node
    (transformer.py: masked 4-head attention + FFN + MaskedLN on the B200 kernels), sum over slots, Linear -> BN."""

    def __init__(self, nhid, nlayer, flavour="alchemy"):
        super().__init__()
        if flavour == "zinc":  # allocated by the reference, never run (core/sign_net.py:54; quirk v)
            self.pos_encoder = MaskedMLP(1, nhid, nlayer=2)
        if nlayer > 0:
            from .transformer import TransformerEncoderLayer
            self.transformer_layers = nn.ModuleList(TransformerEncoderLayer(nhid, n_head=4) for _ in range(nlayer))
        else:
            self.transformer_layers = nn.ModuleList()
        self.out = nn.Sequential(nn.Linear(nhid, nhid, bias=False), nn.BatchNorm1d(nhid))
        self.nhid = nhid

    def reset_parameters(self):
        for layer in self.transformer_layers:
            if hasattr(layer, "reset_parameters"):
                layer.reset_parameters()
        for layer in self.out:  # (the reference iterates an undefined name here, quirk iv)
            if hasattr(layer, "reset_parameters"):
                layer.reset_parameters()

    def forward_rows(self, x_rows, pos_rows, slots):
        """x_rows [2, R, ld] (phi(+v), phi(-v)); pos_rows [R, ld] or None -> [N, nhid]."""
        d = self.nhid
        if len(self.transformer_layers) > 0:
            from .transformer import set_transformer_rows
            x = set_transformer_rows(self, x_rows, pos_rows, slots)
        else:
            x = slot_sum(x_rows, slots, d)
            if pos_rows is not None:
                x = x + slot_sum(pos_rows.unsqueeze(0), slots, d)
        x = linear(x, self.out[0].weight, None, pad4(d))
        x = batch_norm_act(x, self.out[1], self.training, relu=False)
        return x[:, :d] if x.shape[1] != d else x


class SignNet(nn.Module):
    """n x k eigenvectors => n x n_hid, sign invariant and permutation equivariant (sign_net.py:74-118)."""

    def __init__(self, n_hid, nl_phi, nl_rho=2, ignore_eigval=False, flavour="alchemy"):
        super().__init__()
        self.phi = GNN3d(1, n_hid, nl_phi, gnn_type="MaskedGINConv", flavour=flavour)
        self.rho = SetTransformer(n_hid, nl_rho, flavour=flavour)
        self.flavour = flavour
        self.ignore_eigval = ignore_eigval if flavour == "alchemy" else True
        self.n_hid = n_hid
        if flavour == "alchemy":
            if not ignore_eigval:
                self.eigen_encoder = MaskedMLP(1, n_hid, nlayer=2)
        else:  # GINESignNetPyG allocates both, runs neither into the output (core/sign_net.py:90-91,111-112; quirk v)
            self.eigen_encoder1 = MaskedMLP(1, n_hid, nlayer=1)
            self.eigen_encoder2 = MaskedMLP(1, n_hid, nlayer=2)

    def reset_parameters(self):
        self.phi.reset_parameters()
        self.rho.reset_parameters()
        for name in ("eigen_encoder", "eigen_encoder1", "eigen_encoder2"):
            if hasattr(self, name):
                getattr(self, name).reset_parameters()

    def forward(self, data=None, edge_index=None, eigvecs=None, batch=None, edge_attr=None, eigvals=None,
                num_graphs=None, graph_index=None):
        """forward(data) as in the reference, or forward(x, edge_index, eigvecs[N,k], batch, edge_attr, eigvals[N,k])
        (first positional then is the unused node-feature tensor x)."""
        if edge_index is None:
            gi = graph_index or getattr(data, "_b200_graph_index", None) or GraphIndex(
                data.edge_index, data.batch, getattr(data, "num_graphs", None))
            gi.wait_ready()
            try:
                data._b200_graph_index = gi
            except Exception:
                pass
            k = None
            ev_ragged, ev_dense, evals = data.eigen_vectors, None, getattr(data, "eigen_values", None)
        else:
            gi = graph_index or GraphIndex(edge_index, batch, num_graphs)
            ev_ragged, ev_dense, evals = None, eigvecs, eigvals
            k = int(eigvecs.shape[1])
        d = self.n_hid
        if k is None:
            # the reference's k is N_max of the batch (transform.py:31); every graph then has k_b = n_b valid slots
            slots = gi.slots_all(pad4(d))
            k = slots.k
        else:
            slots = gi.slots(k, True, pad4(d))
        x0 = build_phi_input(gi, slots, ev_ragged, ev_dense)
        x_rows, slots = self.phi.forward_rows(x0, gi, k, True)
        pos_rows = None
        if not self.ignore_eigval:
            e0 = torch.empty(slots.R, dtype=torch.float32, device=gi.device)
            if ev_dense is not None:
                if evals is None:
                    raise ValueError("eigvals [N, k] required unless ignore_eigval=True")
                tmp = torch.empty(2, slots.R, dtype=torch.float32, device=gi.device)
                ed = evals.contiguous()
                _call("sb_phi_input_dense", _p(ed), ed.shape[1], _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr),
                      gi.N, slots.k, 1, slots.R, _p(tmp))
                e0 = tmp[0]
            else:
                _call("sb_slot_eigval", _p(evals.contiguous()), _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr),
                      gi.N, slots.k, 1, _p(e0))
            pos_rows = self.eigen_encoder(e0.unsqueeze(1))
        return self.rho.forward_rows(x_rows, pos_rows, slots)


class SignNetGNN(nn.Module):
    """sign_net.SignNetGNN (:120-132): SignNet positional features feeding the GINE predictor."""

    def __init__(self, node_feat, edge_feat, n_hid, n_out, nl_signnet, nl_gnn, nl_rho=4, ignore_eigval=False,
                 gnn_type="GINEConv", flavour="alchemy"):
        super().__init__()
        from .model import GNN
        if flavour == "alchemy":   # the reference ignores its nl_rho argument and always builds 4 (quirk iii)
            self.sign_net = SignNet(n_hid, nl_signnet, nl_rho=4, ignore_eigval=ignore_eigval, flavour=flavour)
        else:                      # GINESignNetPyG/core/sign_net.py:125
            self.sign_net = SignNet(n_hid, nl_signnet, nl_rho=1, flavour=flavour)
        self.gnn = GNN(node_feat, edge_feat, n_hid, n_out, nlayer=nl_gnn, gnn_type=gnn_type,
                       max_num_values=6 if flavour == "alchemy" else 500)

    def reset_parameters(self):
        self.sign_net.reset_parameters()
        self.gnn.reset_parameters()

    def forward(self, data=None, edge_index=None, eigvecs=None, batch=None, edge_attr=None, eigvals=None,
                num_graphs=None):
        if edge_index is None:
            gi = getattr(data, "_b200_graph_index", None) or GraphIndex(data.edge_index, data.batch,
                                                                        getattr(data, "num_graphs", None))
            gi.wait_ready()
            pos = self.sign_net(data, graph_index=gi)
            return self.gnn(data, pos, graph_index=gi)
        gi = GraphIndex(edge_index, batch, num_graphs)
        pos = self.sign_net(data, edge_index, eigvecs, batch, edge_attr, eigvals, num_graphs, graph_index=gi)
        return self.gnn.forward_tensors(data, edge_index, edge_attr, batch, pos, gi)

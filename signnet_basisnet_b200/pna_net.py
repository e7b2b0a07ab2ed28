"""DGL-flavour PNA predictor consuming the sign-invariant positional encoding — the base model of
`configs/pna/PNA_ZINC_LapPE_signinv_GIN{,_mask}.json`.

Mirrors GraphPrediction/nets/ZINC_graph_regression/pna_net.py:18-167 (`PNANet`, the `pe_init='lap_pe'` / no-LSPE,
`gru=False` path), layers/pna_layer.py:16-153 (`PNATower`, `PNALayer`) and the pieces of layers/pna_utils.py they use
(`FCLayer`, `MLP`, aggregators :12-31, scalers :73-84): same `net_params` keys, same state_dict keys
(`layers.{l}.towers.{t}.{pretrans_h,posttrans_h}.fully_connected.0.linear.*`, `layers.{l}.towers.{t}.batchnorm_h.*`,
`layers.{l}.mixing_network_h.linear.*`, `embedding_{h,p,e}`, `MLP_layer.FC_layers.*`, `sign_inv_net.*`),
`forward(g, h, p, e, snorm_n) -> (scores, g)`.  Built for what every shipped configuration selects: aggregators
"mean max min std", scalers "identity amplification attenuation", pretrans_layers = posttrans_layers = 1,
divide_input = True (divide_input = False cannot run in the reference: pna_layer.py:146 passes one argument too many).

Mapping onto kernels.  A tower's pre-transformation Linear(cat[h_src, h_dst, e]) is linear in its three blocks, and the
towers act on disjoint channel slices, so one layer is
    U = h blockdiag(W_src)^T,  V = h blockdiag(W_dst)^T,  Q = e cat(W_e)^T + b        three sb_linear_fwd launches
    Z = [h_t | 4 aggregators x 3 scalers of (U[src] + V[dst] + Q)]_t                   sb_pna_agg_fwd (csrc/pna.cu)
    Y = Z blockdiag(W_post)^T + b;  Y *= snorm_n;  BatchNorm (the towers' statistics are per channel, hence one pass)
    out = h + LeakyReLU(Y W_mix^T + b)                                                  sb_linear_fwd, sb_leaky_relu
STATUS: oracle side pinned against the reference class (oracle/restate.pna_net, fixture tests/golden/dgl_pna_net.pt);
csrc/pna.cu is checked by CPU emulation (tests/test_cpu_emulation_pna.py); GPU parity (kernel vs oracle, PNANet vs the
reference fixture): tests/test_gpu_pna.py, green on the B200.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import counted_call as _call, ptr as _p
from .deepsigns import _graph_index, get_sign_inv_net
from .functional import BatchNormActFn, add_rows, linear
from .gin_net import MLPReadout
from .layout import pad4
from .model import EmbeddingSumFn, SegmentPoolFn

AGGREGATORS = "mean max min std"
SCALERS = "identity amplification attenuation"


class PnaAggFn(torch.autograd.Function):
    """(U, V [N, ld], Q [E, ld], h [N, ldh]) -> Z [N, pad4(13 C)] (tower-major, see csrc/pna.cu)."""

    @staticmethod
    def forward(ctx, U, V, Q, h, gi, C, tin, avg_log):
        U, V, Q, h = (t.contiguous() for t in (U, V, Q, h))
        N, ld = U.shape
        if Q.shape[0] != gi.E or N != gi.N or Q.shape[1] != ld or V.shape != U.shape:
            raise ValueError("PNA aggregate: node / edge tensors do not match the graph")
        ldz = pad4(13 * C)
        Z = torch.empty(N, ldz, dtype=torch.float32, device=U.device)
        _call("sb_pna_agg_fwd", _p(U), _p(V), _p(Q), _p(h), _p(gi.in_ptr), _p(gi.in_src), _p(gi.in_eid), N, C, tin, ld,
              h.shape[1], ldz, float(avg_log), _p(Z))
        ctx.save_for_backward(U, V, Q)
        ctx.cfg = (gi, C, tin, float(avg_log), h.shape[1], ldz)
        return Z

    @staticmethod
    def backward(ctx, dZ):
        U, V, Q = ctx.saved_tensors
        gi, C, tin, avg_log, ldh, ldz = ctx.cfg
        N, ld = U.shape
        dev = U.device
        dZ = dZ.contiguous()
        dU, dV = (torch.empty(N, ld, dtype=torch.float32, device=dev) for _ in range(2))
        dQ = torch.empty(gi.E, ld, dtype=torch.float32, device=dev)
        dh = torch.empty(N, ldh, dtype=torch.float32, device=dev)
        _call("sb_pna_agg_bwd", _p(dZ), _p(U), _p(V), _p(Q), _p(gi.in_ptr), _p(gi.in_src), _p(gi.in_eid), _p(gi.out_ptr),
              _p(gi.out_eid), N, C, tin, ld, ldh, ldz, avg_log, _p(dU), _p(dV), _p(dQ), _p(dh))
        return dU, dV, dQ, dh, None, None, None, None


class RowScaleFn(torch.autograd.Function):
    """out[r, :] = x[r, :] * s[r] (graph normalisation, pna_layer.py:73-74); s carries no gradient."""

    @staticmethod
    def forward(ctx, x, s):
        x, s = x.contiguous(), s.reshape(-1).contiguous().to(torch.float32)
        out = torch.empty_like(x)
        _call("sb_row_scale", _p(x), _p(s), x.shape[0], x.shape[1], _p(out))
        ctx.save_for_backward(s)
        return out

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        g = g.contiguous()
        out = torch.empty_like(g)
        _call("sb_row_scale", _p(g), _p(s), g.shape[0], g.shape[1], _p(out))
        return out, None


class LeakyReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, slope):
        x = x.contiguous()
        out = torch.empty_like(x)
        _call("sb_leaky_relu", None, _p(x), x.numel(), float(slope), _p(out))
        ctx.save_for_backward(x)
        ctx.slope = float(slope)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous()
        out = torch.empty_like(g)
        _call("sb_leaky_relu", _p(g), _p(x), x.numel(), ctx.slope, _p(out))
        return out, None


class FCLayer(nn.Module):
    """pna_utils.py FCLayer restricted to what PNALayer builds: Linear (+ activation applied by the caller), xavier init
    with gain 1 / in_size and zero bias (pna_utils.py reset_parameters)."""

    def __init__(self, in_size, out_size, activation="relu", bias=True):
        super().__init__()
        self.in_size, self.out_size, self.activation = in_size, out_size, activation
        self.linear = nn.Linear(in_size, out_size, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.linear.weight, 1 / self.in_size)
        if self.linear.bias is not None:
            self.linear.bias.data.zero_()


class MLP(nn.Module):
    """pna_utils.py MLP with layers = 1: a single FCLayer without activation (`fully_connected.0`)."""

    def __init__(self, in_size, hidden_size, out_size, layers, mid_activation="relu", last_activation="none"):
        super().__init__()
        if layers != 1:
            raise NotImplementedError("pretrans_layers / posttrans_layers = 1 in every shipped PNA configuration")
        self.fully_connected = nn.ModuleList([FCLayer(in_size, out_size, activation=last_activation)])


class PNATower(nn.Module):
    def __init__(self, in_dim, out_dim, edge_features, edge_dim):
        super().__init__()
        self.batchnorm_h = nn.BatchNorm1d(out_dim)
        self.pretrans_h = MLP(2 * in_dim + (edge_dim if edge_features else 0), in_dim, in_dim, 1)
        self.posttrans_h = MLP(13 * in_dim, out_dim, out_dim, 1)


class PNALayer(nn.Module):
    def __init__(self, in_dim, out_dim, aggregators, scalers, avg_d, dropout, graph_norm, batch_norm, towers=1,
                 pretrans_layers=1, posttrans_layers=1, divide_input=True, residual=False, edge_features=False, edge_dim=0):
        super().__init__()
        if aggregators.split() != AGGREGATORS.split() or scalers.split() != SCALERS.split():
            raise NotImplementedError(f"PNA on the B200 path is built for aggregators '{AGGREGATORS}' and scalers '{SCALERS}'")
        if pretrans_layers != 1 or posttrans_layers != 1 or dropout != 0 or not divide_input or not edge_features:
            raise NotImplementedError("PNALayer: only the shipped form (1-layer pre/post transformations, dropout 0, "
                                      "divide_input, edge features) is built")
        assert in_dim % towers == 0 and out_dim % towers == 0 and avg_d is not None
        self.in_dim, self.out_dim, self.n_towers = in_dim, out_dim, towers
        self.tin, self.tout = in_dim // towers, out_dim // towers
        self.avg_log = float(avg_d["log"])
        self.graph_norm, self.batch_norm = graph_norm, batch_norm
        self.residual = residual and in_dim == out_dim
        self.towers = nn.ModuleList([PNATower(self.tin, self.tout, edge_features, edge_dim) for _ in range(towers)])
        self.mixing_network_h = FCLayer(out_dim, out_dim, activation="LeakyReLU")

    def forward_rows(self, gi, h, e, snorm_n):
        """h [N, pad4(in_dim)], e [E, pad4(edge_dim)], snorm_n [N, 1] -> [N, pad4(out_dim)]."""
        tin = self.tin
        pre_w = [t.pretrans_h.fully_connected[0].linear.weight for t in self.towers]
        pre_b = torch.cat([t.pretrans_h.fully_connected[0].linear.bias for t in self.towers])
        ld = pad4(self.in_dim)
        U = linear(h, torch.block_diag(*[w[:, :tin] for w in pre_w]), None, ld)
        V = linear(h, torch.block_diag(*[w[:, tin:2 * tin] for w in pre_w]), None, ld)
        Q = linear(e, torch.cat([w[:, 2 * tin:] for w in pre_w], dim=0), pre_b, ld)
        Z = PnaAggFn.apply(U, V, Q, h, gi, self.in_dim, tin, self.avg_log)
        post_w = torch.block_diag(*[t.posttrans_h.fully_connected[0].linear.weight for t in self.towers])
        post_b = torch.cat([t.posttrans_h.fully_connected[0].linear.bias for t in self.towers])
        y = linear(Z, post_w, post_b, pad4(self.out_dim))
        if self.graph_norm:
            y = RowScaleFn.apply(y, snorm_n)
        if self.batch_norm:
            y = self._tower_batch_norm(y)
        mix = self.mixing_network_h.linear
        out = LeakyReluFn.apply(linear(y, mix.weight, mix.bias, pad4(self.out_dim)), 0.01)
        return add_rows(h, out) if self.residual else out

    def _tower_batch_norm(self, y):
        """The towers' BatchNorm1d modules act on disjoint channel slices and BatchNorm is per channel: one pass over the
        concatenated parameters, then every tower's running buffers get their slice back."""
        bns = [t.batchnorm_h for t in self.towers]
        gamma, beta = torch.cat([b.weight for b in bns]), torch.cat([b.bias for b in bns])
        rm, rv = torch.cat([b.running_mean for b in bns]), torch.cat([b.running_var for b in bns])
        out = BatchNormActFn.apply(y, gamma, beta, None, rm, rv, self.training, False, self.out_dim, 1)
        if self.training:
            with torch.no_grad():
                for t, b in enumerate(bns):
                    b.running_mean.copy_(rm[t * self.tout:(t + 1) * self.tout])
                    b.running_var.copy_(rv[t * self.tout:(t + 1) * self.tout])
                    b.num_batches_tracked += 1
        return out

    def forward(self, g, h, p, e, snorm_n):
        return self.forward_rows(_graph_index(g, h.device), h, e, snorm_n)[:, :self.out_dim], None


class PNANet(nn.Module):
    def __init__(self, net_params):
        super().__init__()
        hidden_dim, out_dim = net_params["hidden_dim"], net_params["out_dim"]
        n_layers = net_params["L"]
        self.readout = net_params["readout"]
        self.graph_norm = net_params["graph_norm"]
        self.batch_norm = net_params["batch_norm"]
        self.residual = net_params["residual"]
        self.aggregators, self.scalers = net_params["aggregators"], net_params["scalers"]
        self.avg_d = net_params["avg_d"]
        self.towers = net_params["towers"]
        self.divide_input_first = net_params["divide_input_first"]
        self.divide_input_last = net_params["divide_input_last"]
        self.edge_feat = net_params["edge_feat"]
        edge_dim = net_params["edge_dim"]
        self.gru_enable = net_params["gru"]
        self.device = net_params["device"]
        self.pe_init = net_params["pe_init"]
        self.lap_method = net_params["lap_method"]
        self.lap_lspe = net_params["lap_lspe"]
        self.use_lapeig_loss = net_params["use_lapeig_loss"]
        self.lambda_loss, self.alpha_loss = net_params["lambda_loss"], net_params["alpha_loss"]
        self.pos_enc_dim = net_params["pos_enc_dim"]
        if self.pe_init != "lap_pe" or self.lap_lspe:
            raise NotImplementedError("PNANet on the B200 path is the `pe_init='lap_pe'`, no-LSPE predictor")
        if self.gru_enable or not self.edge_feat or self.use_lapeig_loss:
            raise NotImplementedError("PNANet: gru / edge_feat=False / lapeig loss are not built (no shipped sign_inv "
                                      "configuration selects them)")
        if net_params["in_feat_dropout"] != 0 or net_params["dropout"] != 0:
            raise NotImplementedError("dropout is 0.0 in every shipped sign_inv configuration")
        self.embedding_p = nn.Linear(self.pos_enc_dim, hidden_dim)
        self.embedding_h = nn.Embedding(net_params["num_atom_type"], hidden_dim)
        self.embedding_e = nn.Embedding(net_params["num_bond_type"], edge_dim)
        mk = lambda o, div: PNALayer(in_dim=hidden_dim, out_dim=o, dropout=0.0, graph_norm=self.graph_norm,
                                     batch_norm=self.batch_norm, residual=self.residual, aggregators=self.aggregators,
                                     scalers=self.scalers, avg_d=self.avg_d, towers=self.towers, edge_features=self.edge_feat,
                                     edge_dim=edge_dim, divide_input=div,
                                     pretrans_layers=net_params["pretrans_layers"],
                                     posttrans_layers=net_params["posttrans_layers"])
        self.layers = nn.ModuleList([mk(hidden_dim, self.divide_input_first) for _ in range(n_layers - 1)]
                                    + [mk(out_dim, self.divide_input_last)])
        self.MLP_layer = MLPReadout(out_dim, 1)
        self.hidden_dim, self.out_dim = hidden_dim, out_dim
        self.g = None
        if self.lap_method == "sign_inv":
            self.sign_inv_net = get_sign_inv_net(net_params)

    def forward(self, g, h, p, e, snorm_n):
        if not (torch.is_tensor(h) and h.is_cuda):
            raise ValueError("PNANet inputs must be CUDA tensors (no CPU fallback)")
        gi = _graph_index(g, h.device)
        hd = self.hidden_dim
        x = EmbeddingSumFn.apply(h.to(torch.int64), self.embedding_h.weight)                  # [N, pad4(hidden)]
        pp = linear(p.reshape(p.shape[0], -1).contiguous(), self.embedding_p.weight, self.embedding_p.bias, pad4(hd))
        x = add_rows(x, pp)                                                                    # h = h + p (pna_net.py:124)
        ee = EmbeddingSumFn.apply(e.reshape(-1).to(torch.int64), self.embedding_e.weight)      # [E, pad4(edge_dim)]
        for layer in self.layers:
            x = layer.forward_rows(gi, x, ee, snorm_n)
        if self.readout == "max":
            raise NotImplementedError("max readout is not built (no shipped configuration selects it)")
        hg = SegmentPoolFn.apply(x, gi, self.out_dim, self.readout != "sum")
        self.g = g
        return self.MLP_layer(hg), g

    def loss(self, scores, targets):
        return torch.nn.functional.l1_loss(scores, targets)   # pna_net.py:170-178 (task loss)

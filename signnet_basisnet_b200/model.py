"""Downstream GNN predictor of the PyG trees (GINE stack + add-pool + output MLP) on the B200 kernels.

Mirrors Alchemy/sign_net/model.py:9-64, model_utils/pyg_gnn_wrapper.py:7-28 (GINConv, GINEConv) and
model_utils/elements.py:11-69 (Identity, DiscreteEncoder, MLP): same constructor arguments, same state_dict keys.
GAT/GCN/SimplifiedPNA wrappers are out of scope (never selected: main_alchemy.py:35 gnn_type='GINEConv').
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import counted_call as _call, ptr as _p
from .functional import BN_EPS, BN_MOMENTUM, batch_norm_act, linear, linear_fwd, linear_wgrad, wgrad_workspace
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



class GineStackFn(torch.autograd.Function):
    """The whole GINE layer loop of GNN.forward (model.py:44-57) as ONE autograd Function: the launch sequence the
    modules below issue one entry point at a time (edge encoder -> sb_gine_agg_fwd -> Linear -> BN+ReLU -> Linear ->
    BN+ReLU+residual, and the backward of each) is handed to sb_gine_stack_fwd / sb_gine_stack_bwd as a host table of
    device pointers (csrc/gine_stack.cu).  Same kernels, same order, same numbers; ~600 C-ABI calls and their autograd
    nodes / allocations become two calls and a handful of allocations for the reference's 16-layer Alchemy predictor.

    Parameter slots per layer: 0 We | table_0, 1 bn_e.weight|None, 2 bn_e.bias|None, 3 eps, 4 W0, 5 bn0.weight, 6 bn0.bias,
    7 W1, 8 bn.weight, 9 bn.bias, 10..12 table_1..3 | None."""

    PER_LAYER = 13

    @staticmethod
    def _geom(cfg):
        gi, d = cfg["gi"], cfg["d"]
        ld = pad4(d)
        al = lambda n: (n + 63) // 64 * 64
        nN, nE = al(gi.N * ld), al(max(gi.E, 1) * ld)
        return ld, nN, nE, 5 * nN + 2 * nE          # per layer: X_out, A, H, Hn, Y | Ee, e

    @staticmethod
    def _graph_tables(cfg, edge_attr, flags):
        gi = cfg["gi"]
        gp = np.array([gi.in_ptr.data_ptr(), gi.in_src.data_ptr(), gi.in_eid.data_ptr(), gi.out_ptr.data_ptr(),
                       gi.out_dst.data_ptr(), gi.out_eid.data_ptr(), gi.edge_index.data_ptr(), edge_attr.data_ptr(),
                       _p(flags) or 0], dtype=np.int64)
        gn = np.array([gi.N, gi.E, edge_attr.stride(0) if edge_attr.dim() == 2 else 1, cfg["d"], pad4(cfg["d"]),
                       cfg["nfe"], cfg["F"], cfg["V"]], dtype=np.int64)
        return gp, gn

    @staticmethod
    def forward(ctx, x0, edge_attr, cfg, *params):
        gi, L, training, nfe, F = cfg["gi"], cfg["L"], cfg["training"], cfg["nfe"], cfg["F"]
        ld, nN, nE, per_layer = GineStackFn._geom(cfg)
        N, dev = gi.N, x0.device
        x0, edge_attr = x0.contiguous(), edge_attr.contiguous()
        acts = torch.empty(L * per_layer, dtype=torch.float32, device=dev)
        vec32 = torch.empty(L, 6, ld, dtype=torch.float32, device=dev)      # ae, ce, a0, c0, a1, c1
        vec64 = torch.empty(L, 3, 2, ld, dtype=torch.float64, device=dev)   # mre, mr0, mr1  ([2, d] each)
        stats = torch.zeros(L, 3, 2, ld, dtype=torch.float64, device=dev) if training else None
        flags = torch.zeros(1, dtype=torch.int32, device=dev) if F else None
        table = np.zeros((L, 40), dtype=np.int64)
        a0, v32, v64 = acts.data_ptr(), vec32.data_ptr(), vec64.data_ptr()
        st = stats.data_ptr() if training else 0
        X_in = x0.data_ptr()
        P_ = GineStackFn.PER_LAYER
        for l in range(L):
            enc, bne_w, bne_b, eps, W0, g0, b0, W1, g1, b1, t1, t2, t3 = params[l * P_:(l + 1) * P_]
            rme, rve, rm0, rv0, rm1, rv1 = cfg["buffers"][l]
            o = a0 + 4 * l * per_layer
            Xo, A, H, Hn, Y = (o + 4 * i * nN for i in range(5))
            Ee, e = o + 4 * 5 * nN, o + 4 * (5 * nN + nE)
            s_ = [st + 8 * (l * 3 + i) * 2 * ld if training else 0 for i in range(3)]
            a32 = [v32 + 4 * (l * 6 + i) * ld for i in range(6)]
            m64 = [v64 + 8 * (l * 3 + i) * 2 * ld for i in range(3)]
            row = [X_in, Xo, A, H, Hn, Y, Ee if nfe else 0, e,
                   enc.data_ptr() if nfe else 0, _p(bne_w) or 0, _p(bne_b) or 0, _p(rme) or 0, _p(rve) or 0,
                   eps.data_ptr(), W0.data_ptr(), g0.data_ptr(), b0.data_ptr(), _p(rm0) or 0, _p(rv0) or 0,
                   W1.data_ptr(), g1.data_ptr(), b1.data_ptr(), _p(rm1) or 0, _p(rv1) or 0,
                   s_[0] if nfe else 0, s_[1], s_[2],
                   a32[0], a32[1], m64[0], a32[2], a32[3], m64[1], a32[4], a32[5], m64[2]]
            if F:
                row += [t.data_ptr() for t in (enc, t1, t2, t3)[:F]]
            table[l, :len(row)] = row
            X_in = Xo
        gp, gn = GineStackFn._graph_tables(cfg, edge_attr, flags)
        _call("sb_gine_stack_fwd", table.ctypes.data, L, gp.ctypes.data, gn.ctypes.data, int(training), BN_MOMENTUM, BN_EPS)
        out = acts[(L - 1) * per_layer:(L - 1) * per_layer + N * ld].view(N, ld)
        ctx.cfg, ctx.n_params = cfg, len(params)
        ctx.save_for_backward(x0, edge_attr, acts, vec32, vec64, *[p for p in params if p is not None])
        ctx.param_none = [p is None for p in params]
        return out

    @staticmethod
    def backward(ctx, gout):
        cfg = ctx.cfg
        gi, L, d, training, nfe, F, V = cfg["gi"], cfg["L"], cfg["d"], cfg["training"], cfg["nfe"], cfg["F"], cfg["V"]
        ld, nN, nE, per_layer = GineStackFn._geom(cfg)
        x0, edge_attr, acts, vec32, vec64, *rest = ctx.saved_tensors
        it = iter(rest)
        params = [None if none else next(it) for none in ctx.param_none]
        N, dev = gi.N, gout.device
        G = torch.empty(2, nN, dtype=torch.float32, device=dev)            # residual-stream gradient, ping-pong
        G[0, :N * ld].view(N, ld).copy_(gout)
        scr = torch.empty(4 * nN + nE, dtype=torch.float32, device=dev)    # dY, dH, dA, dx | de
        # fp64 statistics arena: one zero-initialised [2, ld] region per BatchNorm of the stack (3 per layer), then coef
        red = torch.zeros(2 * 3 * L + 3, ld, dtype=torch.float64, device=dev)
        deps64 = torch.zeros(L, dtype=torch.float64, device=dev)
        al = lambda n: (n + 63) // 64 * 64
        # one flat block for the parameter gradients; per layer: encoder (We, bn_e.w, bn_e.b | tables), W0, g0, b0, W1, g1, b1
        enc_sizes = [nfe * d, d, d] if nfe else [V * d] * F
        sizes = enc_sizes + [d * d, d, d, d * d, d, d]
        gflat = torch.empty(L * sum(al(n) for n in sizes), dtype=torch.float32, device=dev)
        table = np.zeros((L, 44), dtype=np.int64)
        a0, v32, v64, gb, eb = acts.data_ptr(), vec32.data_ptr(), vec64.data_ptr(), gflat.data_ptr(), deps64.data_ptr()
        grads = [None] * ctx.n_params
        goff, X_in = 0, x0.data_ptr()
        P_ = GineStackFn.PER_LAYER
        for l in range(L):
            base = l * P_
            enc, bne_w, _, eps, W0, g0, _, W1, g1, _, _, _, _ = params[base:base + P_]
            o = a0 + 4 * l * per_layer
            Xo, A, H, Hn, Y = (o + 4 * i * nN for i in range(5))
            Ee, e = o + 4 * 5 * nN, o + 4 * (5 * nN + nE)
            a32 = [v32 + 4 * (l * 6 + i) * ld for i in range(6)]
            m64 = [v64 + 8 * (l * 3 + i) * 2 * ld for i in range(3)]
            ptrs, views = [], []
            for n_ in sizes:
                ptrs.append(gb + 4 * goff)
                views.append(gflat[goff:goff + n_])
                goff += al(n_)
            ne = len(enc_sizes)
            if nfe:
                dWe, dge, dbe, dtabs = ptrs[0], ptrs[1], ptrs[2], []
                grads[base + 0], grads[base + 1], grads[base + 2] = views[0].view(d, nfe), views[1], views[2]
            else:
                dWe = dge = dbe = 0
                dtabs = ptrs[:F]
                grads[base + 0] = views[0].view(V, d)
                for f in range(1, F):
                    grads[base + 9 + f] = views[f].view(V, d)
            dW0, dg0, db0, dW1, dg1, db1 = ptrs[ne:]
            rv = views[ne:]
            grads[base + 4], grads[base + 5], grads[base + 6] = rv[0].view(d, d), rv[1], rv[2]
            grads[base + 7], grads[base + 8], grads[base + 9] = rv[3].view(d, d), rv[4], rv[5]
            row = [X_in, A, H, Hn, Y, Ee if nfe else 0, e, a32[0], a32[1], m64[0], a32[2], a32[3], m64[1], a32[4], a32[5],
                   m64[2], enc.data_ptr() if nfe else 0, _p(bne_w) or 0, eps.data_ptr(), W0.data_ptr(), g0.data_ptr(),
                   W1.data_ptr(), g1.data_ptr(), dWe, dge, dbe, eb + 8 * l, dW0, dg0, db0, dW1, dg1, db1] + dtabs
            table[l, :len(row)] = row
            X_in = Xo
        ews = None
        if F:
            ews = torch.empty(int(_lib.lib().sb_embedding_bwd_workspace_floats(V, d)), dtype=torch.float32, device=dev)
        s0 = scr.data_ptr()
        sc = np.array([G.data_ptr(), G.data_ptr() + 4 * nN, s0, s0 + 4 * nN, s0 + 8 * nN, s0 + 12 * nN, s0 + 16 * nN,
                       red.data_ptr(), red.data_ptr() + 8 * 2 * 3 * L * ld, wgrad_workspace(dev).data_ptr(), _p(ews) or 0],
                      dtype=np.int64)
        gp, gn = GineStackFn._graph_tables(cfg, edge_attr, None)
        _call("sb_gine_stack_bwd", table.ctypes.data, L, gp.ctypes.data, gn.ctypes.data, sc.ctypes.data, int(training))
        deps32 = deps64.to(torch.float32)
        for l in range(L):
            grads[l * P_ + 3] = deps32[l:l + 1]
        gx0 = G[L & 1, :N * ld].view(N, ld)
        return (gx0, None, None, *grads)


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

    def _stack_forward(self, x, edge_attr, gi):
        """The layer loop through sb_gine_stack_* (GineStackFn); None = take the per-module path below (profiling pass,
        SB_GINE_PER_CALL=1, or a configuration the stack driver does not cover)."""
        if _lib._profile is not None or os.environ.get("SB_GINE_PER_CALL") == "1" or not (self.bn and self.res):
            return None
        L, d = len(self.convs), self.nhid
        if L == 0 or x.shape[1] != pad4(d) or gi.E == 0:
            return None
        discrete = isinstance(self.edge_encoders[0], DiscreteEncoder)
        if discrete:
            if edge_attr.is_floating_point() or not edge_attr.is_cuda or edge_attr.dtype != torch.int64:
                return None
            F = 1 if edge_attr.dim() == 1 else edge_attr.shape[1]
            if F > 4:
                return None
            nfe, V = 0, self.edge_encoders[0].embeddings[0].num_embeddings
        else:
            if not (edge_attr.is_floating_point() and edge_attr.is_cuda and edge_attr.dim() == 2):
                return None
            F, V, nfe = 0, 0, edge_attr.shape[1]
            if nfe != self.edge_encoders[0].layers[0].in_features:
                return None
        params, buffers, bns = [], [], []
        for enc, conv, norm in zip(self.edge_encoders, self.convs, self.norms):
            l0, l1 = conv.nn.layers
            bn0 = conv.nn.norms[0]
            if l1.bias is not None or l0.bias is not None:
                return None
            if discrete:
                tabs = [emb.weight for emb in enc.embeddings[:F]]
                params += [tabs[0], None, None]
                rme = rve = None
            else:
                bne = enc.norms[0]
                params += [enc.layers[0].weight, bne.weight, bne.bias]
                rme, rve = bne.running_mean, bne.running_var
                bns.append(bne)
            params += [conv.layer.eps, l0.weight, bn0.weight, bn0.bias, l1.weight, norm.weight, norm.bias]
            params += (tabs[1:] + [None] * 3)[:3] if discrete else [None, None, None]
            buffers.append((rme, rve, bn0.running_mean, bn0.running_var, norm.running_mean, norm.running_var))
            bns += [bn0, norm]
        if self.training:
            counters = [bn.num_batches_tracked for bn in bns
                        if bn.track_running_stats and bn.num_batches_tracked is not None]
            if counters:   # one multi-tensor launch instead of one tiny kernel per BatchNorm
                torch._foreach_add_(counters, 1)
        cfg = dict(gi=gi, L=L, d=d, training=self.training, nfe=nfe, F=F, V=V, buffers=buffers)
        return GineStackFn.apply(x, edge_attr, cfg, *params)

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
        stacked = self._stack_forward(x, edge_attr, gi)
        if stacked is not None:
            x = stacked
            x = SegmentPoolFn.apply(x, gi, d, self.pooling == "mean")
            x = self.output_encoder(x)
            return x[:, :self.nout] if x.shape[1] != self.nout else x
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

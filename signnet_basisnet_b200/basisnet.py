"""The single-graph LearningFilters tree on the B200 kernels: BasisNet's IGN phi (`IGN2to1`, `IGNBasisInv`) and the
DeepSets SignNet (`SignPlus(EqDeepSetsEncoder)`).

Mirrors LearningFilters/ign.py:9-39 (IGN2to1), :88-128 (layer_2_to_1), :174-214 (layer_1_to_1), :344-374 / :404-417
(contractions) and LearningFilters/signbasisnet.py:23-41 (IGNBasisInv): same constructor arguments and state_dict keys
(`equi_layers.{i}.coeffs|bias`, `bns.{0..3}.*`, `fc1.*`, `fc2.*`; the reference registers coeffs/bias as parameters only
when built with device='cpu' - quirk vi of SURVEY.md - they are always parameters here; `bns.3` is allocated and never
applied, as in the reference; `num_layers` is ignored, as in the reference).

Activations live as rows [b * n, C] (eigenspace-major, node-minor) instead of the reference's [b, C, n].  The 2->1
contractions are computed from the eigenvector blocks V_e [n, mult] (`forward_factors`) so the n x n projectors the
reference builds in training.py:59-61 never exist; `forward(proj)` accepts the reference's materialised [b, 1, n, n]
input and reduces it in two coalesced passes.  Everything after the contractions is channel-wise dense work on the
existing Linear / BatchNorm / segment kernels of libsignnet_b200.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ._lib import counted_call as _call, ptr as _p
from .functional import batch_norm_act, linear
from .layout import pad4
from .model import SegmentPoolFn

OPS_LD = 8  # the five contractions, padded to two float4


class _Segments:
    """b eigenspaces of n nodes each as a 'batch' for the segment kernels."""

    def __init__(self, b, n, device):
        self.B, self.N = b, b * n
        self.graph_ptr = (torch.arange(b + 1, device=device, dtype=torch.int64) * n).to(torch.int32)
        self.batch = torch.arange(b, device=device, dtype=torch.int64).repeat_interleave(n)


class _SegBiasActFn(torch.autograd.Function):
    """out[e, i, :] = act(t[e, i, :] + u[e, :]) on [b, n, ld] rows: the broadcast half of a 1->1 equivariant / DeepSets
    layer (act = ReLU or identity)."""

    @staticmethod
    def forward(ctx, t, u, seg, n, C, relu):
        t, u = t.contiguous(), u.contiguous()
        ld = t.shape[1]
        ones = torch.ones(seg.B, C, dtype=torch.float32, device=t.device)
        uc = u[:, :C].contiguous()
        out = torch.empty_like(t)
        _call("sb_affine_act_res", _p(t), _p(ones), _p(uc), None, _p(out), ld, n, seg.B, C, int(relu))
        ctx.save_for_backward(out if relu else None)
        ctx.cfg = (seg, C, u.shape, relu)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        seg, C, ushape, relu = ctx.cfg
        g = g.contiguous()
        gt = g
        if relu:
            gt = torch.empty_like(g)
            _call("sb_relu_bwd", _p(g), _p(out), _p(gt), g.numel())
        gu = torch.zeros(ushape, dtype=torch.float32, device=g.device)
        _call("sb_segment_pool_fwd", _p(gt), gt.stride(0), _p(seg.graph_ptr), seg.B, C, 0, _p(gu), gu.stride(0))
        return gt, gu, None, None, None, None


class _EquiLayer(nn.Module):
    """Parameter holder with the reference's names and init (ign.py:105-112 / :191-198)."""

    def __init__(self, input_depth, output_depth, basis_dimension):
        super().__init__()
        self.input_depth, self.output_depth, self.basis_dimension = input_depth, output_depth, basis_dimension
        self.coeffs = nn.Parameter(torch.randn(input_depth, output_depth, basis_dimension) * math.sqrt(2.0)
                                   / (input_depth + output_depth))
        self.bias = nn.Parameter(torch.zeros(1, output_depth, 1))


class IGN2to1(nn.Module):
    """batch x 1 x n x n projectors (or eigenvector blocks) -> batch x out_channels x n."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers=1, device="cuda", use_bn=True):
        super().__init__()
        if in_channels != 1:
            raise ValueError("IGN2to1 on the B200 path supports in_channels=1 (the only use: signbasisnet.py:33)")
        self.use_bn = use_bn
        self.hidden_channels, self.out_channels = hidden_channels, out_channels
        self.equi_layers = nn.ModuleList([_EquiLayer(in_channels, hidden_channels, 5),
                                          _EquiLayer(hidden_channels, hidden_channels, 2),
                                          _EquiLayer(hidden_channels, hidden_channels, 2)])
        if use_bn:
            self.bns = nn.ModuleList([nn.BatchNorm1d(hidden_channels) for _ in range(4)])
        self.fc1 = nn.Linear(hidden_channels, hidden_channels)
        self.fc2 = nn.Linear(hidden_channels, out_channels)

    # ------------------------------------------------------------------------------------------------ contractions
    @staticmethod
    def ops_from_factors(V, col0, mult):
        """V [n, K] eigenvectors (columns), col0 int32 [b] first column of each eigenspace -> ops rows [b * n, 8]."""
        if not (V.is_cuda and V.dtype == torch.float32):
            raise ValueError("V must be a CUDA float32 tensor (no CPU fallback on the BasisNet path)")
        V = V if V.stride(1) == 1 else V.contiguous()
        n, b = V.shape[0], int(col0.numel())
        ops = torch.empty(b * n, OPS_LD, dtype=torch.float32, device=V.device)
        _call("sb_ign2to1_ops_factors", _p(V), V.stride(0), n, _p(col0.to(torch.int32).contiguous()), b, int(mult),
              _p(ops), OPS_LD)
        return ops

    @staticmethod
    def ops_from_projectors(P):
        """P [b, 1, n, n] (the reference's input, training.py:59-61) -> ops rows [b * n, 8]."""
        if not (P.is_cuda and P.dtype == torch.float32):
            raise ValueError("projectors must be a CUDA float32 tensor (no CPU fallback on the BasisNet path)")
        b, n = P.shape[0], P.shape[-1]
        P = P.reshape(b, n, n).contiguous()
        ops = torch.empty(b * n, OPS_LD, dtype=torch.float32, device=P.device)
        ws = torch.empty(2 * max(b, 1), dtype=torch.float64, device=P.device)
        _call("sb_ign2to1_ops_projectors", _p(P), n, b, _p(ops), OPS_LD, _p(ws))
        return ops

    # ----------------------------------------------------------------------------------------------------- network
    def _bn(self, x, i):
        return batch_norm_act(x, self.bns[i], self.training, relu=False) if self.use_bn else x

    def forward_rows(self, ops, b, n):
        """ops rows [b * n, 8] -> rows [b * n, pad4(out_channels)] (eigenspace-major)."""
        S = self.hidden_channels
        seg = _Segments(b, n, ops.device)
        l0 = self.equi_layers[0]
        # 2->1 layer: einsum('dsb,ndbi->nsi') with d = 1 is a Linear over the five contractions (ign.py:117-128)
        x = linear(ops, l0.coeffs[0], l0.bias.reshape(-1), pad4(S), relu=True)
        x = self._bn(x, 0)
        for i in (1, 2):
            lyr = self.equi_layers[i]
            # 1->1 layer (ign.py:203-214): identity term + mean-over-the-set term + bias
            t = linear(x, lyr.coeffs[:, :, 0].t(), None, pad4(S))
            xm = SegmentPoolFn.apply(x, seg, S, True)
            u = linear(xm, lyr.coeffs[:, :, 1].t(), lyr.bias.reshape(-1), pad4(S))
            x = _SegBiasActFn.apply(t, u, seg, n, S, True)
            x = self._bn(x, i)
        x = linear(x, self.fc1.weight, self.fc1.bias, pad4(S), relu=True)
        return linear(x, self.fc2.weight, self.fc2.bias, pad4(self.out_channels))

    def _to_reference_layout(self, rows, b, n):
        return rows[:, :self.out_channels].reshape(b, n, self.out_channels).transpose(1, 2)

    def forward(self, x):
        """x: projectors [b, 1, n, n] as in the reference -> [b, out_channels, n]."""
        b, n = x.shape[0], x.shape[-1]
        return self._to_reference_layout(self.forward_rows(self.ops_from_projectors(x), b, n), b, n)

    def forward_factors(self, V, col0, mult):
        """V [n, K] eigenvector columns, col0 [b] first column of each eigenspace of multiplicity `mult`."""
        b, n = int(col0.numel()), V.shape[0]
        return self._to_reference_layout(self.forward_rows(self.ops_from_factors(V, col0, mult), b, n), b, n)


class IGNBasisInv(nn.Module):
    """IGN based basis invariant neural network: one IGN2to1 per multiplicity (signbasisnet.py:23-41)."""

    def __init__(self, mult_lst, in_channels, hidden_channels=16, num_layers=2):
        super().__init__()
        self.encs = nn.ModuleList()
        self.mult_to_idx = {}
        for idx, mult in enumerate(mult_lst):
            self.encs.append(IGN2to1(1, hidden_channels, int(mult), num_layers=num_layers))
            self.mult_to_idx[int(mult)] = idx

    def forward(self, proj, mult):
        return self.encs[self.mult_to_idx[int(mult)]](proj)

    def forward_factors(self, V, col0, mult):
        return self.encs[self.mult_to_idx[int(mult)]].forward_factors(V, col0, mult)


class EqDeepSetsEncoder(nn.Module):
    """Equivariant DeepSets encoder, `* x set size x feature size` (LearningFilters/models.py:58-113): phi and rho of the
    single-graph SignNet (training.py:207-218).  Layer: act(lin1(x) + lin2(mean over the set)), BatchNorm1d with
    track_running_stats=False (always batch statistics).  Same constructor and state_dict keys (lins1/lins2/bns)."""

    def __init__(self, in_channels, hidden_channels=32, out_channels=1, num_layers=3, use_bn=False, use_ln=False,
                 dropout=0.0, activation="relu"):
        super().__init__()
        if use_ln or activation != "relu":
            raise NotImplementedError("EqDeepSetsEncoder on the B200 path: use_ln / non-ReLU activations are not built "
                                      "(no reference configuration selects them)")
        dims = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        self.lins1 = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers))
        self.lins2 = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers))
        if use_bn:
            self.bns = nn.ModuleList(nn.BatchNorm1d(hidden_channels, track_running_stats=False)
                                     for _ in range(num_layers - 1))
        self.use_bn, self.use_ln, self.dropout = use_bn, use_ln, dropout

    def forward(self, x, *args):
        if not (x.is_cuda and x.dtype == torch.float32):
            raise ValueError("x must be a CUDA float32 tensor (no CPU fallback)")
        if x.dim() not in (2, 3):
            raise ValueError("invalid x dimension")
        if self.dropout > 0 and self.training:
            raise NotImplementedError("dropout > 0 is not built (every reference configuration uses 0.0)")
        lead = x.shape[0] if x.dim() == 3 else 1
        n, cin = x.shape[-2], x.shape[-1]
        seg = _Segments(lead, n, x.device)
        rows = x.reshape(lead * n, cin).contiguous()
        L = len(self.lins1)
        for i in range(L):
            l1, l2 = self.lins1[i], self.lins2[i]
            C = l1.out_features
            t = linear(rows, l1.weight, l1.bias, pad4(C))
            xm = SegmentPoolFn.apply(rows, seg, l1.in_features, True)
            u = linear(xm, l2.weight, l2.bias, pad4(C))
            last = i == L - 1
            rows = _SegBiasActFn.apply(t, u, seg, n, C, not last)
            if not last and self.use_bn:
                rows = batch_norm_act(rows, self.bns[i], self.training, relu=False)
        out = rows[:, :self.lins1[-1].out_features]
        return out.reshape(lead, n, -1) if x.dim() == 3 else out


class SignPlus(nn.Module):
    """model(v) + model(-v); `x` (not negated) is concatenated on the feature dim (signbasisnet.py:11-20)."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, v, *args, x=None):
        if x is None:
            return self.model(v) + self.model(-v)
        return self.model(torch.cat((v, x), dim=-1)) + self.model(torch.cat((-v, x), dim=-1))


def eigenspace_groups(eigvals: torch.Tensor, decimals: int = 5):
    """Eigenvalue grouping of LearningFilters/training.py:47-62 (round to `decimals`, unique with counts): returns
    {multiplicity: int32 tensor of first-column indices}, eigenspaces in ascending eigenvalue order."""
    r = torch.round(eigvals * 10 ** decimals) / (10 ** decimals)
    _, counts = r.unique(return_counts=True)
    stops = torch.cumsum(counts, 0)
    starts = stops - counts
    out = {}
    for m in counts.unique().tolist():
        out[int(m)] = starts[counts == m].to(torch.int32)
    return out

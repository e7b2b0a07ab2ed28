"""Thin tensor-level wrappers over the C ABI plus the autograd Functions of the dense ([M, d]) pieces of the path.

torch is plumbing here (allocation, autograd bookkeeping, streams); every arithmetic step is a kernel of
libsignnet_b200.  Internal activations are fp32 [rows, ld] with ld = d rounded up to 4 floats and zero padding.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import counted_call as _call, ptr as _p
from .layout import pad4

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

_ws_cache: dict = {}


def _require_cuda(t, name):
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32):
        raise ValueError(f"{name} must be a CUDA float32 tensor (no CPU fallback on the SignNet hot path)")


def wgrad_workspace(device):
    key = (device.type, device.index)
    if key not in _ws_cache:
        n = int(_lib.lib().sb_linear_wgrad_workspace_floats())
        _ws_cache[key] = torch.empty(n, dtype=torch.float32, device=device)
    return _ws_cache[key]


# ------------------------------------------------------------------------------------------------ raw kernel wrappers
def linear_fwd(x, ldx, W, w_rs, w_cs, bias, y, ldy, R, G, K, N, pro=0, pa=None, pc=None, relu=False, stats=None,
               accumulate=False):
    _call("sb_linear_fwd", _p(x), ldx, _p(W), w_rs, w_cs, _p(bias), _p(y), ldy, R, G, K, N, pro, _p(pa), _p(pc),
          int(relu), _p(stats), int(accumulate))


def linear_wgrad(gy, ldg, x, ldx, R, G, N, K, dW, rs, cs, db=None, pro=0, pa=None, pc=None, accumulate=False):
    _call("sb_linear_wgrad", _p(gy), ldg, _p(x), ldx, R, G, N, K, pro, _p(pa), _p(pc), _p(dW), rs, cs, _p(db),
          int(accumulate), _p(wgrad_workspace(gy.device)))


def bn_finalize(stats, M, G, C, gamma, beta, rmean, rvar, training, device):
    """-> (a, c) fp32 [G, C] with BN(x) = a*x + c, and mean_rstd fp64 [2, G, C] for the backward; updates the running
    buffers in place when training."""
    out = torch.empty(2, G, C, dtype=torch.float32, device=device)
    mr = torch.empty(2, G, C, dtype=torch.float64, device=device)
    _call("sb_bn_finalize", _p(stats), M, G, C, _p(gamma), _p(beta), _p(rmean), _p(rvar), BN_MOMENTUM, BN_EPS,
          int(training), _p(out[0]), _p(out[1]), _p(mr))
    return out[0], out[1], mr


def bn_backward(gout, y, a, c, mr, gamma, ld, R, G, C, relu, training, dz_out):
    """dz_out <- gradient w.r.t. the BatchNorm input y, given gout = dL/d act(BN(y)).  Returns (dgamma, dbeta).
    Two streaming passes: the reduction reads (gout, y) and writes nothing; the second pass recomputes the ReLU mask
    from y and writes dz_out (which may alias gout)."""
    dev = y.device
    stats = torch.zeros(G, 2, C, dtype=torch.float64, device=dev)
    _call("sb_bn_bwd_reduce", _p(gout), _p(y), _p(a), _p(c), _p(mr), None, ld, R, G, C, int(relu), _p(stats))
    dgb = torch.empty(2, C, dtype=torch.float32, device=dev)
    _call("sb_bn_apply_bwd", _p(gout), _p(y), _p(stats), _p(mr), _p(a) if relu else None, _p(c) if relu else None,
          _p(gamma), R, int(training), _p(dz_out), _p(dgb[0]), _p(dgb[1]), ld, R, G, C)
    return dgb[0], dgb[1]


# ------------------------------------------------------------------------------------------------- dense autograd ops
class LinearFn(torch.autograd.Function):
    """y[M, ldy] = x[M, :K] @ W[N, K]^T + b.  x may be any row stride (>= K); y is zero padded to ldy."""

    @staticmethod
    def forward(ctx, x, W, b, ldy, relu=False):
        _require_cuda(x, "x")
        x = x if x.stride(-1) == 1 and x.dim() == 2 else x.contiguous()
        W = W.contiguous()
        M, K, N = x.shape[0], W.shape[1], W.shape[0]
        if x.shape[1] < K:
            raise ValueError(f"linear: input has {x.shape[1]} columns, weight expects {K}")
        ldy = N if ldy is None else ldy
        y = torch.empty(M, ldy, dtype=torch.float32, device=x.device)
        linear_fwd(x, x.stride(0), W, K, 1, b, y, ldy, M, 1, K, N, relu=relu)
        ctx.save_for_backward(x, W, y if relu else None)
        ctx.has_bias = b is not None
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, y = ctx.saved_tensors
        gy = gy.contiguous()
        M, K, N = x.shape[0], W.shape[1], W.shape[0]
        if ctx.relu:
            gm = torch.empty_like(gy)
            _call("sb_relu_bwd", _p(gy), _p(y), _p(gm), gy.numel())
            gy = gm
        gx = gW = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty(M, x.shape[1], dtype=torch.float32, device=x.device)
            linear_fwd(gy, gy.stride(0), W, 1, K, None, gx, x.shape[1], M, 1, N, K)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gW = torch.empty_like(W)
            gb = torch.empty(N, dtype=torch.float32, device=x.device) if ctx.has_bias else None
            linear_wgrad(gy, gy.stride(0), x, x.stride(0), M, 1, N, K, gW, K, 1, gb)
        return gx, gW, gb, None, None


def linear(x, W, b=None, ldy=None, relu=False):
    return LinearFn.apply(x, W, b, ldy, relu)


class AddFn(torch.autograd.Function):
    """out = a + b on padded [.., ld] rows."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        ld = a.shape[-1]
        out = torch.empty_like(a)
        _call("sb_affine_act_res", _p(a), None, None, _p(b), _p(out), ld, a.numel() // ld, 1, ld, 0)
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def add_rows(a, b):
    return AddFn.apply(a, b)


class BatchNormActFn(torch.autograd.Function):
    """out = act(BN(x[:, :C])) (+ res) on [M, ld] rows; nn.BatchNorm1d numerics (biased var, eps 1e-5, momentum .1)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, res, rmean, rvar, training, relu, C, G=1):
        """x [M, ld] (G = 1) or [G, M, ld]: statistics are kept per leading group (the two sign passes)."""
        _require_cuda(x, "x")
        x = x.contiguous()
        M, ld = x.shape[-2], x.shape[-1]
        if ld % 4 != 0:
            raise ValueError("batch_norm_act expects a padded activation (ld % 4 == 0)")
        dev = x.device
        ac = torch.empty(2, G, C, dtype=torch.float32, device=dev)
        a, c = ac[0], ac[1]
        # fp64 [G,2,C] scratch of the streaming path followed by mean_rstd [2,G,C]; ONE C-ABI call (one kernel when small)
        f64 = torch.empty(2, 2, G, C, dtype=torch.float64, device=dev)
        mr = f64[1]
        out = torch.empty_like(x)
        _call("sb_bn_act_fwd", _p(x), ld, M, G, C, _p(gamma), _p(beta), _p(rmean), _p(rvar), BN_MOMENTUM, BN_EPS,
              int(training), int(relu), _p(res), _p(out), _p(f64[0]), _p(a), _p(c), _p(mr))
        ctx.save_for_backward(x, a, c, mr, gamma)
        ctx.cfg = (training, relu, C, res is not None, G)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, a, c, mr, gamma = ctx.saved_tensors
        training, relu, C, has_res, G = ctx.cfg
        gout = gout.contiguous()
        M, ld = x.shape[-2], x.shape[-1]
        gx = torch.empty_like(x)
        dgb = torch.empty(2, C, dtype=torch.float32, device=x.device)
        f64 = torch.empty(5, G, C, dtype=torch.float64, device=x.device)   # stats [G,2,C] | coef [3,G,C]
        _call("sb_bn_act_bwd", _p(gout), _p(x), _p(a), _p(c), _p(mr), _p(gamma), ld, M, G, C, int(relu), int(training),
              _p(gx), _p(dgb[0]), _p(dgb[1]), _p(f64[:2]), _p(f64[2:]))
        dgamma, dbeta = dgb[0], dgb[1]
        return gx, dgamma, dbeta, (gout if has_res else None), None, None, None, None, None, None


def batch_norm_act(x, bn: torch.nn.BatchNorm1d, training, relu=True, res=None, C=None, G=1):
    C = bn.num_features if C is None else C
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += G
    use_batch = training or not bn.track_running_stats
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    return BatchNormActFn.apply(x, bn.weight, bn.bias, res, rm, rv, use_batch, relu, C, G)


class SlotSumFn(torch.autograd.Function):
    """[S, R, ld] slot rows -> [N, ldo]: sum over eigenvector slots and sign passes (sign_net.py:113,:70)."""

    @staticmethod
    def forward(ctx, x, slots, C, limit_by_n):
        gi = slots.gi
        S, R, ld = x.shape
        ldo = pad4(C)
        out = torch.empty(gi.N, ldo, dtype=torch.float32, device=x.device)
        _call("sb_slot_sum_fwd", _p(x), ld, R, S, _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr), gi.N, slots.k,
              int(slots.masked), int(limit_by_n), _p(out), ldo, C)
        ctx.cfg = (slots, S, R, ld, C, limit_by_n)
        return out

    @staticmethod
    def backward(ctx, gout):
        slots, S, R, ld, C, limit_by_n = ctx.cfg
        gi = slots.gi
        gout = gout.contiguous()
        gx = torch.empty(S, R, ld, dtype=torch.float32, device=gout.device)
        _call("sb_slot_sum_bwd", _p(gout), gout.stride(0), _p(gx), ld, R, S, _p(gi.batch), _p(gi.graph_ptr),
              _p(slots.row_ptr), gi.N, slots.k, int(slots.masked), int(limit_by_n), C)
        return gx, None, None, None


def slot_sum(x, slots, C, limit_by_n=False):
    return SlotSumFn.apply(x, slots, C, limit_by_n)

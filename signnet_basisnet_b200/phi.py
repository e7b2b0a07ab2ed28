"""phi of SignNet — the L-layer masked-GIN stack applied to +v and -v — as ONE autograd Function over the ragged
slot-row layout, hand-scheduled onto the kernels of libsignnet_b200.

Reference semantics (op order in SURVEY.md Appendix A): GNN3d.forward Alchemy/sign_net/sign_net.py:28-44 with
MaskedGINConv / MaskedMLP / MaskedBN (model_utils/masked_layers.py:13-20,54-64,74-84), invoked twice (sign_net.py:113).
Both sign passes run side by side as the leading S=2 dimension of every activation; BatchNorm statistics are kept per
sign and the running buffers are updated +v first, then -v, exactly as two sequential module calls would.

Per layer l (forward):
    A   = (1+eps_l) X_l + sum_nbr X_l                       sb_gin_agg          (TMA-staged tiles, HBM bound)
    H   = A W0^T                  (+ column stats)          sb_linear_fwd
    Y   = relu(bn0(H)) W1^T + b1  (+ column stats)          sb_linear_fwd       (BN+ReLU applied in the prologue)
    X'  = relu(bn_l(Y)) + X_l                               sb_affine_act_res
Backward recomputes the BN/ReLU element-wise pieces from the saved pre-activations (A, H, Y) instead of storing them.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import counted_call as _call, ptr as _p
from .functional import bn_backward, bn_finalize, linear_fwd, linear_wgrad
from .layout import pad4

PARAMS_PER_LAYER = 8  # W0, bn0.weight, bn0.bias, W1, b1 (or None), eps, bn.weight, bn.bias


def gin_agg(x, out, slots, S, ld, eps=None, res=None, dotx=None, dot_out=None, transpose=False,
            force_generic=False):
    gi = slots.gi
    nbr_ptr, nbr_idx, pack = (gi.out_ptr, gi.out_dst, gi.out_pack) if transpose else (gi.in_ptr, gi.in_src, gi.in_pack)
    generic = force_generic or slots.use_generic_agg or ld % 4 != 0
    _call("sb_gin_agg", _p(x), _p(out), _p(res), _p(dotx), _p(dot_out), _p(eps), _p(gi.graph_ptr),
          _p(slots.unit_ptr), _p(slots.unit_desc), _p(pack), _p(slots.row_ptr), _p(nbr_ptr), _p(nbr_idx), slots.R, gi.B,
          slots.k, int(slots.masked), S, ld, max(slots.tile_rows, 1), int(generic))


class PhiStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, cfg, *params):
        """x0 [S, R] (d_in = 1) or [S, R, ld_in]; cfg = dict(slots_in, slots, dims=[(d_in, h, d)...], training,
        buffers=[(rm0, rv0, rm1, rv1)...]).  Returns X_L [S, R, pad4(d)]."""
        slots, training = cfg["slots"], cfg["training"]
        dims, buffers = cfg["dims"], cfg["buffers"]
        L = len(dims)
        S, R = x0.shape[0], slots.R
        dev = x0.device
        saved, vecs = [], []
        X, ld_in = x0, (1 if x0.dim() == 2 else x0.shape[2])
        for l in range(L):
            d_in, h, d = dims[l]
            W0, g0, b0, W1, b1, eps, g1, bb1 = params[l * PARAMS_PER_LAYER:(l + 1) * PARAMS_PER_LAYER]
            rm0, rv0, rm1, rv1 = buffers[l]
            ldh, ldd = pad4(h), pad4(d)
            sl = cfg["slots_in"] if (l == 0 and ld_in % 4 != 0) else slots
            A = torch.empty(S, R, ld_in, dtype=torch.float32, device=dev) if ld_in > 1 else torch.empty(
                S, R, dtype=torch.float32, device=dev)
            gin_agg(X, A, sl, S, ld_in, eps=eps)
            st0 = torch.zeros(S, 2, h, dtype=torch.float64, device=dev) if training else None
            H = torch.empty(S, R, ldh, dtype=torch.float32, device=dev)
            linear_fwd(A, ld_in, W0, d_in, 1, None, H, ldh, R, S, d_in, h, stats=st0)
            a0, c0, mr0 = bn_finalize(st0, R, S, h, g0, b0, rm0, rv0, training, dev)
            st1 = torch.zeros(S, 2, d, dtype=torch.float64, device=dev) if training else None
            Y = torch.empty(S, R, ldd, dtype=torch.float32, device=dev)
            linear_fwd(H, ldh, W1, h, 1, b1, Y, ldd, R, S, h, d, pro=2, pa=a0, pc=c0, stats=st1)
            a1, c1, mr1 = bn_finalize(st1, R, S, d, g1, bb1, rm1, rv1, training, dev)
            Xn = torch.empty(S, R, ldd, dtype=torch.float32, device=dev)
            _call("sb_affine_act_res", _p(Y), _p(a1), _p(c1), _p(X if l > 0 else None), _p(Xn), ldd, R, S, d, 1)
            saved += [X, A, H, Y]
            vecs += [a0, c0, mr0, a1, c1, mr1]
            if cfg.get("capture") is not None:  # test hook: pre-activations + BN affines (activation patterns)
                cfg["capture"].append(dict(H=H, a0=a0, c0=c0, Y=Y, a1=a1, c1=c1, h=h, d=d))
            X, ld_in = Xn, ldd
        ctx.cfg = cfg
        ctx.n_params = len(params)
        ctx.save_for_backward(*saved, *vecs, *[p for p in params if p is not None])
        ctx.param_none = [p is None for p in params]
        return X

    @staticmethod
    def backward(ctx, gout):
        cfg = ctx.cfg
        slots, training, dims = cfg["slots"], cfg["training"], cfg["dims"]
        L = len(dims)
        tensors = list(ctx.saved_tensors)
        saved, vecs = tensors[:4 * L], tensors[4 * L:10 * L]
        it = iter(tensors[10 * L:])
        params = [None if none else next(it) for none in ctx.param_none]
        S, R = gout.shape[0], slots.R
        dev = gout.device
        G = gout.contiguous().clone()  # dL/dX_{l+1}; updated in place down the residual stream
        grads = [None] * ctx.n_params
        for l in reversed(range(L)):
            d_in, h, d = dims[l]
            X, A, H, Y = saved[4 * l:4 * l + 4]
            a0, c0, mr0, a1, c1, mr1 = vecs[6 * l:6 * l + 6]
            W0, g0, b0, W1, b1, eps, g1, bb1 = params[l * PARAMS_PER_LAYER:(l + 1) * PARAMS_PER_LAYER]
            ld_in = 1 if X.dim() == 2 else X.shape[2]
            ldh, ldd = pad4(h), pad4(d)
            base = l * PARAMS_PER_LAYER
            # outer BN + ReLU:  dY
            dY = torch.empty(S, R, ldd, dtype=torch.float32, device=dev)
            grads[base + 6], grads[base + 7] = bn_backward(G, Y, a1, c1, mr1, g1, ldd, R, S, d, True, training, dY)
            # second Linear: dW1, db1 (input recomputed as relu(bn0(H)) in the prologue), dP
            gW1 = torch.empty_like(W1)
            gb1 = torch.empty_like(b1) if b1 is not None else None
            linear_wgrad(dY, ldd, H, ldh, R, S, d, h, gW1, h, 1, gb1, pro=2, pa=a0, pc=c0)
            grads[base + 3], grads[base + 4] = gW1, gb1
            dH = torch.empty(S, R, ldh, dtype=torch.float32, device=dev)
            linear_fwd(dY, ldd, W1, 1, h, None, dH, ldh, R, S, d, h)
            del dY
            # inner BN + ReLU:  dH (in place)
            grads[base + 1], grads[base + 2] = bn_backward(dH, H, a0, c0, mr0, g0, ldh, R, S, h, True, training, dH)
            # first Linear: dW0, dA
            gW0 = torch.empty_like(W0)
            linear_wgrad(dH, ldh, A, ld_in, R, S, h, d_in, gW0, d_in, 1, None)
            grads[base + 0] = gW0
            dA = torch.empty_like(A)
            linear_fwd(dH, ldh, W0, 1, d_in, None, dA, ld_in, R, S, h, d_in)
            del dH
            # aggregate (transposed CSR) + residual stream + d eps
            deps = torch.zeros(1, dtype=torch.float64, device=dev)
            sl = cfg["slots_in"] if (l == 0 and ld_in % 4 != 0) else slots
            if l > 0:
                gin_agg(dA, G, sl, S, ld_in, eps=eps, res=G, dotx=X, dot_out=deps, transpose=True)
            else:
                scratch = torch.empty_like(dA)
                gin_agg(dA, scratch, sl, S, ld_in, eps=eps, dotx=X, dot_out=deps, transpose=True)
            grads[base + 5] = deps.to(torch.float32)
        return (None, None, *grads)

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

Two drivers issue the SAME launch sequence: the default one hands a pointer table to sb_phi_stack_fwd / sb_phi_stack_bwd
(csrc/phi_stack.cu: the layer loop runs in host C++, two C-ABI calls and five allocations per step instead of ~160
calls and ~100 allocations - the step is host-bound at 128 graphs per GPU otherwise); the per-call driver below it is
kept for bench.py's per-entry-point CUDA-event breakdown (`_lib.profile_start()`) and selected with SB_PHI_PER_CALL=1.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from ._lib import counted_call as _call, ptr as _p
from .functional import BN_EPS, BN_MOMENTUM, bn_backward, bn_finalize, linear_fwd, linear_wgrad, wgrad_workspace
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


def _per_call():
    return _lib._profile is not None or os.environ.get("SB_PHI_PER_CALL") == "1"


def _slot_tables(cfg):
    """Host tables of the two slot layouts (sb_phi_stack_* contract, include/signnet_b200.h); cached on the layout."""
    sl, sl_in = cfg["slots"], cfg["slots_in"]
    key = "_phi_tables"
    hit = sl.__dict__.get(key)
    if hit is not None and hit[0] is sl_in:
        return hit[1], hit[2]
    ptrs = np.empty((2, 10), dtype=np.int64)
    ints = np.empty((2, 6), dtype=np.int64)
    for r, s_ in enumerate((sl, sl_in)):
        gi = s_.gi
        ptrs[r] = [gi.graph_ptr.data_ptr(), s_.unit_ptr.data_ptr(), s_.unit_desc.data_ptr(), gi.in_pack.data_ptr(),
                   gi.out_pack.data_ptr(), s_.row_ptr.data_ptr(), gi.in_ptr.data_ptr(), gi.in_src.data_ptr(),
                   gi.out_ptr.data_ptr(), gi.out_dst.data_ptr()]
        ints[r] = [s_.R, gi.B, s_.k, int(s_.masked), max(s_.tile_rows, 1), int(s_.use_generic_agg)]
    sl.__dict__[key] = (sl_in, ptrs, ints)
    return ptrs, ints


def _dp(t):
    return 0 if t is None else t.data_ptr()


def _al(n):
    """Sub-tensor offsets inside the packed blocks are kept 256-byte aligned (vector loads, TMA)."""
    return (n + 63) // 64 * 64


class PhiStackFn(torch.autograd.Function):
    """Default driver: the layer loop runs inside libsignnet_b200 (sb_phi_stack_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x0, cfg, *params):
        if _per_call():
            return _PhiStackPerCallFn.forward(ctx, x0, cfg, *params)
        slots, training = cfg["slots"], cfg["training"]
        dims, buffers = cfg["dims"], cfg["buffers"]
        L, S, R, dev = len(dims), x0.shape[0], slots.R, x0.device
        ld0 = 1 if x0.dim() == 2 else x0.shape[2]
        # one block for every activation the backward needs: per layer A [ld_in], H [ldh], Y [ldd], X_out [ldd]
        widths, cmax = [], 1
        for l, (d_in, h, d) in enumerate(dims):
            ld_in = ld0 if l == 0 else pad4(dims[l - 1][2])
            widths.append((ld_in, pad4(h), pad4(d), pad4(d)))
            cmax = max(cmax, pad4(h), pad4(d))
        rows = S * R
        acts = torch.empty(sum(_al(rows * w_) for w in widths for w_ in w), dtype=torch.float32, device=dev)
        vec32 = torch.empty(L, 4, S, cmax, dtype=torch.float32, device=dev)            # a0, c0, a1, c1
        vec64 = torch.empty(L, 2, 2, S, cmax, dtype=torch.float64, device=dev)         # mr0, mr1  ([2,S,C] each)
        stats = torch.zeros(L, 2, S, 2, cmax, dtype=torch.float64, device=dev) if training else None
        table = np.empty((L, 25), dtype=np.int64)
        dtab = np.empty((L, 4), dtype=np.int32)
        a_base, v32, v64 = acts.data_ptr(), vec32.data_ptr(), vec64.data_ptr()
        st = stats.data_ptr() if training else 0
        off, X_in, layer_off = 0, x0.data_ptr(), []
        for l, (d_in, h, d) in enumerate(dims):
            W0, g0, b0, W1, b1, eps, g1, bb1 = params[l * PARAMS_PER_LAYER:(l + 1) * PARAMS_PER_LAYER]
            rm0, rv0, rm1, rv1 = buffers[l]
            o = [off]
            for w in widths[l]:
                o.append(o[-1] + _al(rows * w))
            layer_off.append(o)
            off = o[-1]
            A, H, Y, Xn = (a_base + 4 * v for v in o[:4])
            q32 = v32 + 4 * l * 4 * S * cmax
            q64 = v64 + 8 * l * 4 * S * cmax
            qst = st + 8 * l * 4 * S * cmax if training else 0
            table[l] = [X_in, A, H, Y, Xn, W0.data_ptr(), g0.data_ptr(), b0.data_ptr(), W1.data_ptr(), _dp(b1),
                        eps.data_ptr(), g1.data_ptr(), bb1.data_ptr(), _dp(rm0), _dp(rv0), _dp(rm1), _dp(rv1),
                        qst, qst + 8 * 2 * S * cmax if training else 0,
                        q32, q32 + 4 * S * cmax, q64, q32 + 8 * S * cmax, q32 + 12 * S * cmax, q64 + 8 * 2 * S * cmax]
            dtab[l] = [d_in, h, d, widths[l][0]]
            X_in = Xn
        sp, si = _slot_tables(cfg)
        _call("sb_phi_stack_fwd", table.ctypes.data, dtab.ctypes.data, L, sp.ctypes.data, si.ctypes.data, S,
              int(training), BN_MOMENTUM, BN_EPS)
        ldd = widths[-1][3]
        out = acts[layer_off[-1][3]:layer_off[-1][3] + rows * ldd].view(S, R, ldd)
        if cfg.get("capture") is not None:  # test hook: pre-activations + BN affines (activation patterns)
            for l, (d_in, h, d) in enumerate(dims):
                o, w = layer_off[l], widths[l]
                cfg["capture"].append(dict(H=acts[o[1]:o[1] + rows * w[1]].view(S, R, w[1]),
                                           Y=acts[o[2]:o[2] + rows * w[2]].view(S, R, w[2]),
                                           a0=vec32[l, 0].reshape(-1)[:S * h].view(S, h), c0=vec32[l, 1].reshape(-1)[:S * h].view(S, h),
                                           a1=vec32[l, 2].reshape(-1)[:S * d].view(S, d), c1=vec32[l, 3].reshape(-1)[:S * d].view(S, d),
                                           h=h, d=d))
        ctx.cfg = cfg
        ctx.n_params = len(params)
        ctx.layout = (layer_off, widths, cmax, ld0)
        ctx.per_call = False
        ctx.save_for_backward(x0, acts, vec32, vec64, *[p for p in params if p is not None])
        ctx.param_none = [p is None for p in params]
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, gout):
        if ctx.per_call:
            return _PhiStackPerCallFn.backward(ctx, gout)
        cfg = ctx.cfg
        slots, training, dims = cfg["slots"], cfg["training"], cfg["dims"]
        layer_off, widths, cmax, ld0 = ctx.layout
        L = len(dims)
        x0, acts, vec32, vec64, *rest = ctx.saved_tensors
        it = iter(rest)
        params = [None if none else next(it) for none in ctx.param_none]
        S, R, dev = gout.shape[0], slots.R, gout.device
        rows = S * R
        G = gout.contiguous().clone()          # dL/dX_{l+1}; updated in place down the residual stream
        wmax = max(max(w) for w in widths)
        scratch_t = torch.empty(4, _al(rows * wmax), dtype=torch.float32, device=dev)       # dY, dH, dA, layer-0 sink
        # fp64 statistics arena, one zero-initialised [S,2,C] region per BatchNorm of the stack (ONE fill per backward)
        red = torch.zeros(2 * L, 2 * S * cmax, dtype=torch.float64, device=dev)
        deps64 = torch.zeros(L, dtype=torch.float64, device=dev)
        # one flat block for every parameter gradient of the stack
        sizes = []
        for l, (d_in, h, d) in enumerate(dims):
            b1 = params[l * PARAMS_PER_LAYER + 4]
            sizes.append((h * d_in, h, h, d * h, d if b1 is not None else 0, d, d))
        gflat = torch.empty(sum(_al(n_) for s_ in sizes for n_ in s_), dtype=torch.float32, device=dev)
        table = np.empty((L, 23), dtype=np.int64)
        dtab = np.empty((L, 4), dtype=np.int32)
        a_base, v32, v64, g_base, e_base = acts.data_ptr(), vec32.data_ptr(), vec64.data_ptr(), gflat.data_ptr(), deps64.data_ptr()
        grads, goff, X = [None] * ctx.n_params, 0, x0.data_ptr()
        for l, (d_in, h, d) in enumerate(dims):
            W0, g0, b0, W1, b1, eps, g1, bb1 = params[l * PARAMS_PER_LAYER:(l + 1) * PARAMS_PER_LAYER]
            o = layer_off[l]
            A, H, Y, Xn = (a_base + 4 * v for v in o[:4])
            q32 = v32 + 4 * l * 4 * S * cmax
            q64 = v64 + 8 * l * 4 * S * cmax
            gp, views = [], []
            for n_, shape in zip(sizes[l], ((h, d_in), (h,), (h,), (d, h), (d,), (d,), (d,))):
                gp.append(g_base + 4 * goff if n_ else 0)
                views.append(gflat[goff:goff + n_].view(shape) if n_ else None)
                goff += _al(n_)
            table[l] = [X, A, H, Y, q32, q32 + 4 * S * cmax, q64, q32 + 8 * S * cmax, q32 + 12 * S * cmax,
                        q64 + 8 * 2 * S * cmax, W0.data_ptr(), g0.data_ptr(), W1.data_ptr(), eps.data_ptr(), g1.data_ptr(),
                        gp[0], gp[1], gp[2], gp[3], gp[4], e_base + 8 * l, gp[5], gp[6]]
            dtab[l] = [d_in, h, d, widths[l][0]]
            base = l * PARAMS_PER_LAYER
            grads[base + 0], grads[base + 1], grads[base + 2], grads[base + 3], grads[base + 4] = views[:5]
            grads[base + 6], grads[base + 7] = views[5], views[6]
            X = Xn
        s0 = scratch_t.data_ptr()
        step = 4 * _al(rows * wmax)
        scr = np.array([G.data_ptr(), s0, s0 + step, s0 + 2 * step, s0 + 3 * step, red.data_ptr(),
                        2 * S * cmax, wgrad_workspace(dev).data_ptr()], dtype=np.int64)
        sp, si = _slot_tables(cfg)
        _call("sb_phi_stack_bwd", table.ctypes.data, dtab.ctypes.data, L, sp.ctypes.data, si.ctypes.data,
              scr.ctypes.data, S, int(training))
        deps32 = deps64.to(torch.float32)
        for l in range(L):
            grads[l * PARAMS_PER_LAYER + 5] = deps32[l:l + 1]
        return (None, None, *grads)


class _PhiStackPerCallFn:
    """Per-call driver (one C-ABI call per kernel): same launch sequence, used while profiling per entry point."""

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
            st0 = torch.zeros(S, 2, h, dtype=torch.float64, device=dev) if training else None
            H = torch.empty(S, R, ldh, dtype=torch.float32, device=dev)
            gi = sl.gi   # aggregate + first Linear: the fused kernel where the shape allows it, else the two kernels
            if not _lib.try_call("sb_gin_linear_fused_fwd", _p(X), _p(A), _p(H), _p(st0), _p(eps), _p(W0), d_in, 1, d_in, h,
                                 ldh, _p(sl.unit_ptr), _p(sl.unit_desc), _p(gi.in_pack), _p(gi.in_ptr), _p(gi.in_src),
                                 sl.R, gi.B, S, ld_in, max(sl.tile_rows, 1), int(sl.use_generic_agg or ld_in % 4 != 0)):
                gin_agg(X, A, sl, S, ld_in, eps=eps)
                linear_fwd(A, ld_in, W0, d_in, 1, None, H, ldh, R, S, d_in, h, stats=st0)
            a0, c0, mr0 = bn_finalize(st0, R, S, h, g0, b0, rm0, rv0, training, dev)
            st1 = torch.zeros(S, 2, d, dtype=torch.float64, device=dev) if training else None
            Y = torch.empty(S, R, ldd, dtype=torch.float32, device=dev)
            linear_fwd(H, ldh, W1, h, 1, b1, Y, ldd, R, S, h, d, pro=2, pa=a0, pc=c0, stats=st1)
            Xn = torch.empty(S, R, ldd, dtype=torch.float32, device=dev)
            ac1 = torch.empty(2, S, d, dtype=torch.float32, device=dev)
            a1, c1 = ac1[0], ac1[1]
            mr1 = torch.empty(2, S, d, dtype=torch.float64, device=dev)
            _call("sb_bn_apply_fwd", _p(Y), _p(st1), R, S, d, _p(g1), _p(bb1), _p(rm1), _p(rv1), BN_MOMENTUM, BN_EPS,
                  int(training), 1, _p(X if l > 0 else None), _p(Xn), ldd, R, _p(a1), _p(c1), _p(mr1))
            saved += [X, A, H, Y]
            vecs += [a0, c0, mr0, a1, c1, mr1]
            if cfg.get("capture") is not None:  # test hook: pre-activations + BN affines (activation patterns)
                cfg["capture"].append(dict(H=H, a0=a0, c0=c0, Y=Y, a1=a1, c1=c1, h=h, d=d))
            X, ld_in = Xn, ldd
        ctx.cfg = cfg
        ctx.n_params = len(params)
        ctx.per_call = True
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

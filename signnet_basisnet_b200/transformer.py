"""rho of the PyG trees — SetTransformer over the eigenvector slots of every node — on the B200 kernels.

Mirrors Alchemy/sign_net/model_utils/transformer_module.py:27-127 (TransformerEncoderLayer, MultiHeadAttention,
ScaledDotProductAttention, PositionwiseFeedForward) and masked_layers.MaskedLN (:22-32) with the reference's
state_dict keys.  Tokens are the valid slot rows of a node, so all `x[~mask] = 0` writes of the reference vanish.

Reference quirk kept: MultiHeadAttention builds ScaledDotProductAttention with its default attn_dropout = 0.1
(transformer_module.py:46,85), i.e. attention probabilities are dropped in training mode even though every other
dropout of the model is 0.  It is reproduced with a counter-based generator inside the kernels, seeded per layer call
from torch's CPU generator (not torch's CUDA Philox stream: the masks differ from the reference's element for element, so
training-mode parity is checked with the dropout set to 0, cf. oracle/restate.py:set_transformer; the statistics - keep
probability, 1/(1-p) scaling, fresh mask per layer and step, same mask in forward and backward - are tested).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import counted_call as _call, ptr as _p
from .functional import add_rows, linear, linear_fwd, linear_wgrad, slot_sum
from .layout import pad4


class AttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, slots, n_head, dk, drop_p, seed):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        gi = slots.gi
        ld = q.shape[1]
        # every slot row belongs to exactly one node: the kernel writes all of o unless there are padding columns
        o = torch.empty_like(q) if n_head * dk == ld else torch.zeros_like(q)
        temp = float(dk) ** 0.5
        _call("sb_attention_fwd", _p(q), _p(k), _p(v), ld, _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr), gi.N,
              slots.k, int(slots.masked), max(slots.kmax, 1), n_head, dk, temp, float(drop_p), int(seed), _p(o))
        ctx.save_for_backward(q, k, v)
        ctx.cfg = (slots, n_head, dk, temp, float(drop_p), int(seed))
        return o

    @staticmethod
    def backward(ctx, go):
        q, k, v = ctx.saved_tensors
        slots, n_head, dk, temp, drop_p, seed = ctx.cfg
        gi = slots.gi
        go = go.contiguous()
        alloc = torch.empty_like if n_head * dk == q.shape[1] else torch.zeros_like
        gq, gk, gv = alloc(q), alloc(q), alloc(q)
        _call("sb_attention_bwd", _p(q), _p(k), _p(v), _p(go), q.shape[1], _p(gi.batch), _p(gi.graph_ptr),
              _p(slots.row_ptr), gi.N, slots.k, int(slots.masked), max(slots.kmax, 1), n_head, dk, temp, drop_p, seed,
              _p(gq), _p(gk), _p(gv))
        return gq, gk, gv, None, None, None, None, None


class LayerNormFn(torch.autograd.Function):
    """y = LN(a + b) * w + beta over the first C columns of padded rows."""

    @staticmethod
    def forward(ctx, a, b, w, beta, C, eps):
        a = a.contiguous()
        b = b.contiguous() if b is not None else None
        R, ld = a.shape
        y = torch.empty_like(a)
        xsum = torch.empty_like(a)
        stat = torch.empty(R, 2, dtype=torch.float32, device=a.device)
        _call("sb_layernorm_fwd", _p(a), _p(b), _p(w), _p(beta), ld, R, C, float(eps), _p(y), _p(xsum), _p(stat))
        ctx.save_for_backward(xsum, stat, w)
        ctx.cfg = (C, b is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        xsum, stat, w = ctx.saved_tensors
        C, has_b = ctx.cfg
        g = g.contiguous()
        R, ld = xsum.shape
        dx = torch.empty_like(xsum)
        dwb = torch.zeros(2, C, dtype=torch.float64, device=g.device)
        _call("sb_layernorm_bwd", _p(g), _p(xsum), _p(stat), _p(w), ld, R, C, _p(dx), _p(dwb))
        dwb = dwb.to(torch.float32)
        return dx, (dx if has_b else None), dwb[0], dwb[1], None, None


class MaskedLN(nn.Module):
    def __init__(self, num_features):
        super().__init__()
        self.ln = nn.LayerNorm(num_features, eps=1e-6)

    def reset_parameters(self):
        self.ln.reset_parameters()

    def forward(self, x_rows, residual=None):
        return LayerNormFn.apply(x_rows, residual, self.ln.weight, self.ln.bias, self.ln.normalized_shape[0],
                                 self.ln.eps)


def _ln_fwd(a, b, w, beta, C, eps):
    R, ld = a.shape
    y, xsum = torch.empty_like(a), torch.empty_like(a)
    stat = torch.empty(R, 2, dtype=torch.float32, device=a.device)
    _call("sb_layernorm_fwd", _p(a), _p(b), _p(w), _p(beta), ld, R, C, float(eps), _p(y), _p(xsum), _p(stat))
    return y, xsum, stat


def _ln_bwd(g, xsum, stat, w, C):
    """-> (dx = gradient w.r.t. BOTH LayerNorm summands, dweight, dbias)."""
    R, ld = xsum.shape
    dx = torch.empty_like(xsum)
    dwb = torch.zeros(2, C, dtype=torch.float64, device=g.device)
    _call("sb_layernorm_bwd", _p(g), _p(xsum), _p(stat), _p(w), ld, R, C, _p(dx), _p(dwb))
    dwb = dwb.to(torch.float32)
    return dx, dwb[0], dwb[1]


def _lin(x, W, b, ldy, relu=False):
    """y [M, ldy] = x[:, :K] W^T (+ b) (relu)."""
    M, (N, K) = x.shape[0], W.shape
    y = torch.empty(M, ldy, dtype=torch.float32, device=x.device)
    linear_fwd(x, x.stride(0), W, K, 1, b, y, ldy, M, 1, K, N, relu=relu)
    return y


def _lin_dx(gy, W, out_ld, into=None):
    """gx [M, out_ld] (+)= gy[:, :N] W: the input gradient of y = x W^T; accumulates into `into` when given."""
    M, (N, K) = gy.shape[0], W.shape
    gx = into if into is not None else torch.empty(M, out_ld, dtype=torch.float32, device=gy.device)
    linear_fwd(gy, gy.stride(0), W, 1, K, None, gx, gx.stride(0), M, 1, N, K, accumulate=into is not None)
    return gx


def _lin_dw(gy, x, W, bias):
    M, (N, K) = gy.shape[0], W.shape
    gW = torch.empty_like(W)
    gb = torch.empty(N, dtype=torch.float32, device=gy.device) if bias else None
    linear_wgrad(gy, gy.stride(0), x, x.stride(0), M, 1, N, K, gW, K, 1, gb)
    return gW, gb


class MHABlockFn(torch.autograd.Function):
    """LN(fc(attention(x Wq^T, x Wk^T, x Wv^T)) + x) as ONE autograd node (transformer_module.py:76-102).  Built from
    separate nodes, x has four consumers and autograd adds their gradients with three activation-sized eager `add`
    launches; here the three projection input-gradients accumulate straight into the LayerNorm's residual gradient
    (sb_linear_fwd with accumulate = 1)."""

    @staticmethod
    def forward(ctx, x, Wq, Wk, Wv, Wfc, lnw, lnb, slots, n_head, dk, d_model, eps, drop_p, seed):
        _require(x)
        x = x.contiguous()
        Wq, Wk, Wv, Wfc = Wq.contiguous(), Wk.contiguous(), Wv.contiguous(), Wfc.contiguous()
        gi, hd = slots.gi, pad4(n_head * dk)
        q, k, v = _lin(x, Wq, None, hd), _lin(x, Wk, None, hd), _lin(x, Wv, None, hd)
        o = torch.empty_like(q) if n_head * dk == hd else torch.zeros_like(q)
        temp = float(dk) ** 0.5
        _call("sb_attention_fwd", _p(q), _p(k), _p(v), hd, _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr), gi.N,
              slots.k, int(slots.masked), max(slots.kmax, 1), n_head, dk, temp, float(drop_p), int(seed), _p(o))
        of = _lin(o, Wfc, None, x.shape[1])
        y, xsum, stat = _ln_fwd(of, x, lnw, lnb, d_model, eps)
        ctx.save_for_backward(x, q, k, v, o, xsum, stat, Wq, Wk, Wv, Wfc, lnw)
        ctx.cfg = (slots, n_head, dk, d_model, temp, float(drop_p), int(seed))
        return y

    @staticmethod
    def backward(ctx, gy):
        x, q, k, v, o, xsum, stat, Wq, Wk, Wv, Wfc, lnw = ctx.saved_tensors
        slots, n_head, dk, d_model, temp, drop_p, seed = ctx.cfg
        gi, hd = slots.gi, q.shape[1]
        dx, dlnw, dlnb = _ln_bwd(gy.contiguous(), xsum, stat, lnw, d_model)   # d(fc output) = d(residual x)
        go = _lin_dx(dx, Wfc, hd)
        gWfc, _ = _lin_dw(dx, o, Wfc, False)
        alloc = torch.empty_like if n_head * dk == hd else torch.zeros_like
        gq, gk, gv = alloc(q), alloc(q), alloc(q)
        _call("sb_attention_bwd", _p(q), _p(k), _p(v), _p(go), hd, _p(gi.batch), _p(gi.graph_ptr), _p(slots.row_ptr),
              gi.N, slots.k, int(slots.masked), max(slots.kmax, 1), n_head, dk, temp, drop_p, seed, _p(gq), _p(gk), _p(gv))
        gWq, _ = _lin_dw(gq, x, Wq, False)
        gWk, _ = _lin_dw(gk, x, Wk, False)
        gWv, _ = _lin_dw(gv, x, Wv, False)
        gx = dx                                   # residual gradient; the three projections accumulate into it
        _lin_dx(gq, Wq, None, into=gx)
        _lin_dx(gk, Wk, None, into=gx)
        _lin_dx(gv, Wv, None, into=gx)
        return gx, gWq, gWk, gWv, gWfc, dlnw, dlnb, None, None, None, None, None, None, None


class FFNBlockFn(torch.autograd.Function):
    """LN(W2 relu(W1 x + b1) + b2 + x) as one autograd node (transformer_module.py:105-127): the input gradient of W1
    accumulates into the LayerNorm's residual gradient instead of an eager add."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, lnw, lnb, d_in, eps):
        _require(x)
        x = x.contiguous()
        W1, W2 = W1.contiguous(), W2.contiguous()
        h = _lin(x, W1, b1, pad4(W1.shape[0]), relu=True)
        h2 = _lin(h, W2, b2, x.shape[1])
        y, xsum, stat = _ln_fwd(h2, x, lnw, lnb, d_in, eps)
        ctx.save_for_backward(x, h, xsum, stat, W1, W2, lnw)
        ctx.cfg = (d_in, b1 is not None, b2 is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, h, xsum, stat, W1, W2, lnw = ctx.saved_tensors
        d_in, has_b1, has_b2 = ctx.cfg
        dx, dlnw, dlnb = _ln_bwd(gy.contiguous(), xsum, stat, lnw, d_in)
        gW2, gb2 = _lin_dw(dx, h, W2, has_b2)
        gh = _lin_dx(dx, W2, h.shape[1])
        _call("sb_relu_bwd", _p(gh), _p(h), _p(gh), gh.numel())     # in place: gh *= [h > 0]
        gW1, gb1 = _lin_dw(gh, x, W1, has_b1)
        gx = dx
        _lin_dx(gh, W1, None, into=gx)
        return gx, gW1, gb1, gW2, gb2, dlnw, dlnb, None, None


def _require(x):
    if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
        raise ValueError("rho expects CUDA float32 slot rows [R, ld] (no CPU fallback on the SignNet hot path)")


class _AttentionDropout(nn.Module):
    """Placeholder mirroring ScaledDotProductAttention.dropout so `.attention.dropout.p` can be set like on the
    reference module."""

    def __init__(self, temperature, attn_dropout=0.1):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)


class MultiHeadAttention(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):
        super().__init__()
        if d_k != d_v:
            raise ValueError("the reference always uses d_k == d_v")
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
        self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
        self.attention = _AttentionDropout(temperature=d_k ** 0.5)
        self.dropout = nn.Dropout(dropout)
        self.norm = MaskedLN(d_model)
        self.d_model = d_model
        self._calls = 0

    def forward(self, x_rows, slots):
        p = self.attention.dropout.p if self.training else 0.0
        self._calls += 1
        seed = 0
        if p > 0.0:
            # one draw per layer call from torch's CPU generator: independent masks per layer, per step and - when the
            # ranks are seeded differently, as DDP scripts do - per rank; reproducible under torch.manual_seed and
            # resumable with torch.get_rng_state() (the seed used to be a function of the call count alone,
            # so all rho layers of a step dropped the same attention entries)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if x_rows.shape[1] == pad4(self.d_model):
            return MHABlockFn.apply(x_rows, self.w_qs.weight, self.w_ks.weight, self.w_vs.weight, self.fc.weight,
                                    self.norm.ln.weight, self.norm.ln.bias, slots, self.n_head, self.d_k, self.d_model,
                                    self.norm.ln.eps, p, seed)
        hd = pad4(self.n_head * self.d_k)
        q = linear(x_rows, self.w_qs.weight, None, hd)
        k = linear(x_rows, self.w_ks.weight, None, hd)
        v = linear(x_rows, self.w_vs.weight, None, hd)
        o = AttentionFn.apply(q, k, v, slots, self.n_head, self.d_k, p, seed)
        o = linear(o, self.fc.weight, None, pad4(self.d_model))
        return self.norm(o, x_rows)


class PositionwiseFeedForward(nn.Module):
    def __init__(self, d_in, d_hid, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_in, d_hid)
        self.w_2 = nn.Linear(d_hid, d_in)
        self.norm = MaskedLN(d_in)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x_rows):
        if x_rows.shape[1] == pad4(self.w_2.out_features):
            return FFNBlockFn.apply(x_rows, self.w_1.weight, self.w_1.bias, self.w_2.weight, self.w_2.bias,
                                    self.norm.ln.weight, self.norm.ln.bias, self.w_2.out_features, self.norm.ln.eps)
        h = linear(x_rows, self.w_1.weight, self.w_1.bias, pad4(self.w_1.out_features), relu=True)
        h = linear(h, self.w_2.weight, self.w_2.bias, pad4(self.w_2.out_features))
        return self.norm(h, x_rows)


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, n_head, dropout=0):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_model // n_head, d_model // n_head, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_model, dropout=dropout)

    def forward(self, x_rows, slots):
        return self.pos_ffn(self.slf_attn(x_rows, slots))


def set_transformer_rows(rho, x_rows, pos_rows, slots):
    """SetTransformer.forward up to the sum over slots (sign_net.py:60-70): x_rows [2, R, ld] -> [N, pad4(d)]."""
    x = add_rows(x_rows[0], x_rows[1])
    if pos_rows is not None:
        x = add_rows(x, pos_rows)
    for layer in rho.transformer_layers:
        x = layer(x, slots)
    return slot_sum(x.unsqueeze(0), slots, rho.nhid)

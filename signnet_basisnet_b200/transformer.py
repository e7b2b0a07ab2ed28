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
from .functional import add_rows, linear, slot_sum
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
        hd = pad4(self.n_head * self.d_k)
        q = linear(x_rows, self.w_qs.weight, None, hd)
        k = linear(x_rows, self.w_ks.weight, None, hd)
        v = linear(x_rows, self.w_vs.weight, None, hd)
        p = self.attention.dropout.p if self.training else 0.0
        self._calls += 1
        seed = 0
        if p > 0.0:
            # one draw per layer call from torch's CPU generator: independent masks per layer, per step and - when the
            # ranks are seeded differently, as DDP scripts do - per rank; reproducible under torch.manual_seed and
            # resumable with torch.get_rng_state() (the seed used to be a function of the call count alone,
            # so all rho layers of a step dropped the same attention entries)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
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

"""rho's self-attention (Alchemy/sign_net/model_utils/transformer_module.py:44-58,76-102) on the tensor-core kernels of
csrc/attention_mma.cu: forward and all three input gradients against a dense fp64 restatement at 1e-5, against the FFMA
kernels they replace (same dropout mask, so the two must agree to rounding), and the shapes at the edges of the fast path
(k_b = 1, 16/17, 32/33, 37 tokens; unmasked fixed k; d_k != 32 falls back)."""
import pytest
import torch

from helpers import slot_row_index
from signnet_basisnet_b200.synth import synth_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dense_reference(q, k, v, w, idx, H, dk):
    """fp64 autograd attention per node over its valid slots; q, k, v, w [R, H*dk] -> o and the gradients of sum(o*w)."""
    q, k, v = (t.double().cpu().requires_grad_(True) for t in (q, k, v))
    valid = idx >= 0                                     # [N, kmax]
    safe = idx.clamp(min=0)
    N, K = idx.shape

    def heads(t):
        return t[safe].view(N, K, H, dk).permute(0, 2, 1, 3)   # [N, H, K, dk]

    s = (heads(q) / dk ** 0.5) @ heads(k).transpose(-1, -2)
    s = s.masked_fill(~valid[:, None, None, :], float("-inf"))
    o = torch.softmax(s, -1) @ heads(v)                  # [N, H, K, dk]
    o = o.permute(0, 2, 1, 3).reshape(N, K, H * dk)
    out = torch.zeros_like(q)
    out = out.index_put((idx[valid],), o[valid])
    (out * w.double().cpu()).sum().backward()
    return out.detach(), q.grad, k.grad, v.grad


def _run(q, k, v, w, sl, H, dk, p, seed, mma):
    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200.transformer import AttentionFn

    old = _lib.lib().sb_set_attention_mma(mma)
    try:
        q, k, v = (t.clone().requires_grad_(True) for t in (q, k, v))
        o = AttentionFn.apply(q, k, v, sl, H, dk, p, seed)
        (o * w).sum().backward()
        torch.cuda.synchronize()
        return o.detach(), q.grad, k.grad, v.grad
    finally:
        _lib.lib().sb_set_attention_mma(old)


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp(min=1e-30))


@pytest.mark.parametrize("shape,B,k,masked", [("zinc", 24, None, True), ("alchemy", 33, None, True),
                                              ("zinc", 9, 8, True), ("zinc", 7, 8, False), ("zinc", 5, 37, False)])
def test_attention_mma_matches_fp64_and_ffma(shape, B, k, masked):
    from signnet_basisnet_b200.layout import GraphIndex

    d = synth_batch(B, shape, seed=40 + B)
    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    sl = gi.slots_all(128) if k is None else gi.slots(k, masked, 128)
    idx = slot_row_index(d.batch, sl.k, sl.masked)
    H, dk = 4, 32
    g = torch.Generator(device=DEV).manual_seed(B)
    q, kk, v, w = (torch.randn(sl.R, H * dk, device=DEV, generator=g) * s for s in (1.5, 1.5, 1.0, 1.0))
    ref = _dense_reference(q, kk, v, w, idx, H, dk)
    got = _run(q, kk, v, w, sl, H, dk, 0.0, 0, 1)
    ffma = _run(q, kk, v, w, sl, H, dk, 0.0, 0, 0)
    for name, a, f, r in zip(("o", "dq", "dk", "dv"), got, ffma, ref):
        assert _rel(a, r) <= 1e-5, (name, _rel(a, r), _rel(f, r))
        assert _rel(a, f) <= 1e-5, (name, "mma vs ffma", _rel(a, f))
    # training-mode dropout: the mask is a function of (seed, node, head, query, key), identical in both kernel families
    got = _run(q, kk, v, w, sl, H, dk, 0.1, 777, 1)
    ffma = _run(q, kk, v, w, sl, H, dk, 0.1, 777, 0)
    for name, a, f in zip(("o", "dq", "dk", "dv"), got, ffma):
        assert _rel(a, f) <= 1e-5, (name, "dropout: mma vs ffma", _rel(a, f))
    assert not torch.equal(got[0], _run(q, kk, v, w, sl, H, dk, 0.0, 0, 1)[0])


def test_attention_token_counts_at_the_tile_edges():
    """Graphs of exactly 1, 2, 8, 9, 16, 17, 24, 25, 32, 33 and 37 nodes (k_b = n_b): every row/key tile boundary of the
    m16n8k8 tiling."""
    from signnet_basisnet_b200.layout import GraphIndex

    sizes = [1, 2, 8, 9, 16, 17, 24, 25, 32, 33, 37, 40]
    batch = torch.cat([torch.full((n,), i, dtype=torch.int64) for i, n in enumerate(sizes)])
    start = torch.tensor([0] + sizes).cumsum(0)
    src = torch.cat([torch.arange(n - 1) + start[i] for i, n in enumerate(sizes)])      # path graphs
    ei = torch.stack([torch.cat([src, src + 1]), torch.cat([src + 1, src])]).to(torch.int64)
    gi = GraphIndex(ei.to(DEV), batch.to(DEV), len(sizes))
    sl = gi.slots_all(128)
    idx = slot_row_index(batch, sl.k, True)
    H, dk = 4, 32
    g = torch.Generator(device=DEV).manual_seed(3)
    q, kk, v, w = (torch.randn(sl.R, H * dk, device=DEV, generator=g) for _ in range(4))
    ref = _dense_reference(q, kk, v, w, idx, H, dk)
    got = _run(q, kk, v, w, sl, H, dk, 0.0, 0, 1)
    for name, a, r in zip(("o", "dq", "dk", "dv"), got, ref):
        assert _rel(a, r) <= 1e-5, (name, _rel(a, r))


def test_other_head_widths_fall_back():
    """d_k = 8 (n_hid 32) is outside the tensor-core path: the FFMA kernels answer, results unchanged by the switch."""
    from signnet_basisnet_b200.layout import GraphIndex

    d = synth_batch(6, "alchemy", seed=33)
    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    sl = gi.slots_all(32)
    g = torch.Generator(device=DEV).manual_seed(1)
    q, k, v, w = (torch.randn(sl.R, 32, device=DEV, generator=g) for _ in range(4))
    a = _run(q, k, v, w, sl, 4, 8, 0.0, 0, 1)
    b = _run(q, k, v, w, sl, 4, 8, 0.0, 0, 0)
    for x, y in zip(a, b):
        assert torch.equal(x, y)

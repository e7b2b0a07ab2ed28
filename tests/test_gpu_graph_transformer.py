"""Sparse graph-Transformer predictor on the GPU against the reference's own output / gradients
(tests/golden/dgl_transformer_net.pt) and the CPU oracle.  The kernels' source is additionally checked on the
CPU by tests/test_cpu_emulation_attention.py and the oracle by tests/test_oracle_vs_reference.py."""
import os

import pytest
import torch

import restate
from helpers import assert_close_rel, assert_grads_close, assert_grads_parity, assert_parity  # noqa: F401
from signnet_basisnet_b200.synth import Data, synth_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _leaf64(sd):
    """fp64 leaf copy of a fixture's state_dict (the arbiter of helpers.assert_parity is the oracle run in fp64)."""
    out = {k: (v.detach().clone().double() if v.is_floating_point() else v.detach().clone()) for k, v in sd.items()}
    for k, v in out.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    return out


def _pe64(d, sd64, prm, masked=True):
    sub = {k[len("sign_inv_net."):]: v for k, v in sd64.items() if k.startswith("sign_inv_net.")}
    x = d.pos_enc.unsqueeze(-1).double()
    if masked:
        return restate.masked_gin_deepsigns(x, d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph, sub,
                                            prm["sign_inv_layers"], prm["pos_enc_dim"]).squeeze(-1)
    return restate.gin_deepsigns(x, d.edge_index[0], d.edge_index[1], sub, prm["sign_inv_layers"],
                                 prm["pos_enc_dim"]).squeeze(-1)


def _g64(sd64, want):
    return {k: sd64[k].grad for k in want}


class _G:
    def __init__(self, d):
        self.src, self.dst, self.n = d.edge_index[0], d.edge_index[1], torch.as_tensor(d.num_nodes_per_graph)

    def edges(self):
        return self.src, self.dst

    def batch_num_nodes(self):
        return self.n


@pytest.mark.parametrize("H,d,scale", [(4, 4, 1.0), (8, 8, 1.0), (2, 5, 3.0)])
def test_edge_attention_kernel_vs_oracle(H, d, scale):
    from signnet_basisnet_b200.graph_transformer_net import EdgeAttentionFn
    from signnet_basisnet_b200.layout import GraphIndex, pad4

    g_ = synth_batch(9, "zinc", seed=35)
    N, E, C = g_.batch.numel(), g_.edge_index.shape[1], H * d
    src, dst = g_.edge_index
    gen = torch.Generator().manual_seed(6)
    Qr, Kr, Vr = (torch.randn(N, C, generator=gen) * scale for _ in range(3))
    Er = torch.randn(E, C, generator=gen) * scale
    w = torch.randn(N, C, generator=gen)

    def oracle(dt):
        q, k, e, v = (t.to(dt).clone().requires_grad_(True) for t in (Qr, Kr, Er, Vr))
        o = restate.sparse_attention(q, k, e, v, src, dst, H)
        (o * w.to(dt)).sum().backward()
        return [t.double() for t in (o.detach(), q.grad, k.grad, e.grad, v.grad)]

    o32, o64 = oracle(torch.float32), oracle(torch.float64)
    tol = [max(2e-5, 3.0 * float((a - b).abs().max())) for a, b in zip(o32, o64)]
    gi = GraphIndex(g_.edge_index.to(DEV), g_.batch.to(DEV), g_.num_graphs)
    ld = pad4(C)
    pad = lambda t: torch.nn.functional.pad(t, (0, ld - C)).to(DEV).requires_grad_(True)
    Q, K, Ef, V = pad(Qr), pad(Kr), pad(Er), pad(Vr)
    out = EdgeAttentionFn.apply(Q, K, Ef, V, gi, H, d)
    assert float((out[:, :C].double().cpu() - o64[0]).abs().max()) <= tol[0]
    (out[:, :C] * w.to(DEV)).sum().backward()
    for name, got, want, t in (("dQ", Q.grad, o64[1], tol[1]), ("dK", K.grad, o64[2], tol[2]), ("dE", Ef.grad, o64[3], tol[3]),
                               ("dV", V.grad, o64[4], tol[4])):
        assert float((got[:, :C].double().cpu() - want).abs().max()) <= t, name


def test_transformer_net_golden(golden_dir):
    from signnet_basisnet_b200.gatedgcn_net import handle_lap
    from signnet_basisnet_b200.graph_transformer_net import TransformerNet

    g = torch.load(os.path.join(golden_dir, "dgl_transformer_net.pt"), weights_only=False)
    d, prm = Data(**g["data"]).to(DEV), dict(g["params"], device=DEV)
    net = TransformerNet(prm).to(DEV).train()
    assert set(net.state_dict()) == set(g["state_dict"])
    net.load_state_dict(g["state_dict"])
    G = _G(d)
    dc, sd64 = d.to("cpu"), _leaf64(g["state_dict"])
    ref64 = restate.transformer_net(dc.x[:, 0], _pe64(dc, sd64, prm, masked=False), dc.edge_attr.reshape(-1),
                                    dc.edge_index[0], dc.edge_index[1], dc.num_nodes_per_graph, sd64, prm["L"],
                                    prm["n_heads"], prm["readout"], prm["pe_aggregate"])
    (ref64 * g["w"].double()).sum().backward()
    pe = handle_lap(net, d.pos_enc, G, DEV)
    out, g_ret = net(G, d.x[:, 0], pe, d.edge_attr.reshape(-1), None)
    assert g_ret is G and out.shape == g["out"].shape
    assert_parity(out, g["out"], ref64, TOL, what="TransformerNet vs reference")
    (out * g["w"].to(DEV)).sum().backward()
    got = {k: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"])
    assert_grads_parity(got, g["grads"], _g64(sd64, g["grads"]), TOL, "TransformerNet vs reference")
    after = net.state_dict()
    for k, v in g["state_dict_after"].items():
        if "running_" in k and k.startswith("layers."):
            assert_parity(after[k], v, sd64[k], TOL, what=k)

"""Sparse graph-Transformer predictor on the GPU against the reference's own output / gradients
(tests/golden/dgl_transformer_net.pt) and the CPU oracle.  csrc/graph_attention.cu and graph_transformer_net.py were
written after the round's GPU budget was spent: these tests run last in the GPU session as non-strict xfail; the kernels' source is checked on the
CPU by tests/test_cpu_emulation_attention.py and the oracle by tests/test_oracle_vs_reference.py."""
import os

import pytest
import torch

import restate
from helpers import assert_close_rel, assert_grads_close
from signnet_basisnet_b200.synth import Data, synth_batch

# Written after round 1's last GPU visit.  The kernels involved are plain streaming kernels (no barriers, no tensor cores:
# nothing that can hang), their source is emulated on the CPU and the module wiring is dry-run, so the tests are allowed
# to run - LAST in the session (file name) and as non-strict xfail: XPASS = parity observed, XFAIL = needs work, never red.
pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="not yet observed on a GPU (XPASS = parity holds)")]
DEV = "cuda"


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
    pe = handle_lap(net, d.pos_enc, G, DEV)
    out, g_ret = net(G, d.x[:, 0], pe, d.edge_attr.reshape(-1), None)
    assert g_ret is G and out.shape == g["out"].shape
    assert_close_rel(out.detach().cpu(), g["out"], 2e-5, what="TransformerNet vs reference")
    (out * g["w"].to(DEV)).sum().backward()
    got = {k: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"])
    assert_grads_close(got, g["grads"], 1e-4, "TransformerNet vs reference")
    after = net.state_dict()
    for k, v in g["state_dict_after"].items():
        if "running_" in k and k.startswith("layers."):
            torch.testing.assert_close(after[k].cpu(), v, rtol=1e-4, atol=1e-5)

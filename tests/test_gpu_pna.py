"""PNA predictor on the GPU against the reference's own output / gradients
(tests/golden/dgl_pna_net.pt) and the CPU oracle.  The kernels' source is additionally
checked on the CPU by tests/test_cpu_emulation_pna.py and the oracle by tests/test_oracle_vs_reference.py."""
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


def test_pna_aggregate_kernel_vs_oracle():
    from signnet_basisnet_b200.layout import GraphIndex, pad4
    from signnet_basisnet_b200.pna_net import PnaAggFn

    C, tin, avg = 20, 4, 1.1
    d = synth_batch(9, "zinc", seed=33)
    N, E = d.batch.numel(), d.edge_index.shape[1]
    src, dst = d.edge_index
    gen = torch.Generator().manual_seed(2)
    Ur, Vr, hr = (torch.randn(N, C, generator=gen) for _ in range(3))
    Qr = torch.randn(E, C, generator=gen)
    kc, rc = [], []
    for t in range(C // tin):
        for s in range(3):
            for a in range(4):
                for j in range(tin):
                    kc.append(t * 13 * tin + tin + (s * 4 + a) * tin + j)
                    rc.append((s * 4 + a) * C + t * tin + j)
    kc, rc = torch.tensor(kc), torch.tensor(rc)
    wz = torch.randn(N, 12 * C, generator=gen)

    def oracle(dt):
        U, V, Q = (t.to(dt).clone().requires_grad_(True) for t in (Ur, Vr, Qr))
        agg = restate.pna_aggregate((U[src] + V[dst]) + Q, dst, N, avg)[:, rc]
        (agg * wz.to(dt)).sum().backward()
        return [t.double() for t in (agg.detach(), U.grad, V.grad, Q.grad)]

    o32, o64 = oracle(torch.float32), oracle(torch.float64)
    tol = [max(5e-5, 2.0 * float((a - b).abs().max())) for a, b in zip(o32, o64)]   # std is ill-conditioned near 0 variance
    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    ld = pad4(C)
    pad = lambda t: torch.nn.functional.pad(t, (0, ld - C)).to(DEV).requires_grad_(True)
    U, V, Q, h = pad(Ur), pad(Vr), pad(Qr), pad(hr)
    Z = PnaAggFn.apply(U, V, Q, h, gi, C, tin, avg)
    assert float((Z[:, kc.to(DEV)].double().cpu() - o64[0]).abs().max()) <= tol[0]
    (Z[:, kc.to(DEV)] * wz.to(DEV)).sum().backward()
    for name, got, want, t in (("dU", U.grad, o64[1], tol[1]), ("dV", V.grad, o64[2], tol[2]), ("dQ", Q.grad, o64[3], tol[3])):
        assert float((got[:, :C].double().cpu() - want).abs().max()) <= t, name


def test_pna_net_golden(golden_dir):
    """PNANet(net_params) with its sign_inv_net on the GPU vs the reference's own output, gradients and BN buffers."""
    from signnet_basisnet_b200.gatedgcn_net import handle_lap
    from signnet_basisnet_b200.pna_net import PNANet

    g = torch.load(os.path.join(golden_dir, "dgl_pna_net.pt"), weights_only=False)
    d, prm = Data(**g["data"]).to(DEV), dict(g["params"], device=DEV)
    net = PNANet(prm).to(DEV).train()
    assert set(net.state_dict()) == set(g["state_dict"])
    net.load_state_dict(g["state_dict"])
    G = _G(d)
    dc, sd64 = d.to("cpu"), _leaf64(g["state_dict"])
    ref64 = restate.pna_net(dc.x[:, 0], _pe64(dc, sd64, prm), dc.edge_attr.reshape(-1), dc.edge_index[0], dc.edge_index[1],
                            dc.num_nodes_per_graph, g["snorm_n"].double(), sd64, prm["L"], prm["towers"],
                            prm["avg_d"]["log"], prm["readout"], prm["divide_input_first"], prm["divide_input_last"])
    (ref64 * g["w"].double()).sum().backward()
    pe = handle_lap(net, d.pos_enc, G, DEV)
    out, g_ret = net(G, d.x[:, 0], pe, d.edge_attr.reshape(-1), g["snorm_n"].to(DEV))
    assert g_ret is G and out.shape == g["out"].shape
    # (the std aggregator is ill-conditioned near zero variance: that is what the fp64 arbiter is for)
    assert_parity(out, g["out"], ref64, TOL, what="PNANet vs reference")
    (out * g["w"].to(DEV)).sum().backward()
    got = {k: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"])
    assert_grads_parity(got, g["grads"], _g64(sd64, g["grads"]), TOL, "PNANet vs reference")
    after = net.state_dict()
    for k, v in g["state_dict_after"].items():
        if "running_" in k and k.startswith("layers."):
            assert_parity(after[k], v, sd64[k], TOL, what=k)

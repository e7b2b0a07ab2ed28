"""GatedGCN predictor + PE baselines on the GPU against the reference's own output / gradients
(tests/golden/dgl_gatedgcn_net.pt) and the CPU oracle.

Parity first observed on a B200 in round 2; these are plain gates now.  The oracle side is pinned by
tests/test_oracle_vs_reference.py / tests/test_oracle_golden.py on the CPU."""
import os
import types

import pytest
import torch

import restate
from helpers import (assert_close_rel, assert_grads_close, assert_grads_parity, assert_parity,  # noqa: F401
                     fp32_noise_samples)
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
    """What the DGL-flavour modules touch on a batched graph: edges() and batch_num_nodes()."""

    def __init__(self, d):
        self.src, self.dst, self.n = d.edge_index[0], d.edge_index[1], torch.as_tensor(d.num_nodes_per_graph)

    def edges(self):
        return self.src, self.dst

    def batch_num_nodes(self):
        return self.n


def test_gated_aggregate_kernel_vs_oracle():
    """sb_gated_agg_fwd/bwd against the plain-torch statement of gatedgcn_layer.py:48-54 (fp64 arbiter)."""
    from signnet_basisnet_b200.gatedgcn_net import GatedAggFn
    from signnet_basisnet_b200.layout import GraphIndex

    d = synth_batch(9, "zinc", seed=31)
    N, E, C = d.batch.numel(), d.edge_index.shape[1], 20
    gen = torch.Generator().manual_seed(2)
    ins = [torch.randn(N, C, generator=gen, dtype=torch.float64) for _ in range(4)] + [torch.randn(E, C, generator=gen, dtype=torch.float64)]
    wh, we = torch.randn(N, C, generator=gen, dtype=torch.float64), torch.randn(E, C, generator=gen, dtype=torch.float64)
    src, dst = d.edge_index
    ref_in = [t.clone().requires_grad_(True) for t in ins]
    Ah, Bh, Dh, Eh, Ce = ref_in
    e_ref = (Dh[src] + Eh[dst]) + Ce
    sg = torch.sigmoid(e_ref)
    h_ref = Ah + torch.zeros(N, C, dtype=torch.float64).index_add(0, dst, Bh[src] * sg) / (
        torch.zeros(N, C, dtype=torch.float64).index_add(0, dst, sg) + 1e-6)
    ((h_ref * wh).sum() + (e_ref * we).sum()).backward()

    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    cu_in = [t.float().to(DEV).requires_grad_(True) for t in ins]
    h, e = GatedAggFn.apply(*cu_in, gi)
    assert_close_rel(h.detach().cpu(), h_ref.detach().float(), 1e-5, what="gated aggregate h")
    assert_close_rel(e.detach().cpu(), e_ref.detach().float(), 1e-5, what="gated aggregate e")
    ((h * wh.float().to(DEV)).sum() + (e * we.float().to(DEV)).sum()).backward()
    for name, a, b in zip("A B D E C".split(), cu_in, ref_in):
        assert_close_rel(a.grad.cpu(), b.grad.float(), 2e-5, what=f"d{name}h")


def test_gatedgcn_net_golden(golden_dir):
    """GatedGCNNet(net_params) with its sign_inv_net on the GPU vs the reference's own output, gradients, BN buffers."""
    from signnet_basisnet_b200.gatedgcn_net import GatedGCNNet, handle_lap

    g = torch.load(os.path.join(golden_dir, "dgl_gatedgcn_net.pt"), weights_only=False)
    d, prm = Data(**g["data"]).to(DEV), dict(g["params"], device=DEV)
    net = GatedGCNNet(prm).to(DEV).train()
    assert set(net.state_dict()) == set(g["state_dict"])
    net.load_state_dict(g["state_dict"])
    G = _G(d)
    dc, sd64 = d.to("cpu"), _leaf64(g["state_dict"])
    ref64 = restate.gatedgcn_net(dc.x[:, 0], _pe64(dc, sd64, prm), dc.edge_attr.reshape(-1), dc.edge_index[0],
                                 dc.edge_index[1], dc.num_nodes_per_graph, sd64, prm["L"], prm["readout"],
                                 prm["edge_feat"], prm["pe_aggregate"])
    (ref64 * g["w"].double()).sum().backward()
    pe = handle_lap(net, d.pos_enc, G, DEV)                                # 'sign_inv', train_ZINC_graph_regression.py:20-25
    out, g_ret = net(G, d.x[:, 0], pe, d.edge_attr.reshape(-1), None)
    assert g_ret is G and out.shape == g["out"].shape
    assert_parity(out, g["out"], ref64, TOL, what="GatedGCNNet vs reference")
    (out * g["w"].to(DEV)).sum().backward()
    got = {k: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"])
    def run(sd_):
        sub = {k[len("sign_inv_net."):]: v for k, v in sd_.items() if k.startswith("sign_inv_net.")}
        pe_ = restate.masked_gin_deepsigns(dc.pos_enc.unsqueeze(-1), dc.edge_index[0], dc.edge_index[1],
                                           dc.num_nodes_per_graph, sub, prm["sign_inv_layers"], prm["pos_enc_dim"]).squeeze(-1)
        o = restate.gatedgcn_net(dc.x[:, 0], pe_, dc.edge_attr.reshape(-1), dc.edge_index[0], dc.edge_index[1],
                                 dc.num_nodes_per_graph, sd_, prm["L"], prm["readout"], prm["edge_feat"], prm["pe_aggregate"])
        (o * g["w"]).sum().backward()
        return {k: sd_[k].grad for k in g["grads"]}

    assert_grads_parity(got, g["grads"], _g64(sd64, g["grads"]), TOL, "GatedGCNNet vs reference",
                        samples=fp32_noise_samples(run, g["state_dict"], 3))
    after = net.state_dict()
    for k, v in g["state_dict_after"].items():
        if "running_" in k and k.startswith("layers."):
            assert_parity(after[k], v, sd64[k], TOL, what=k)


@pytest.mark.parametrize("method", ["sign_flip", "abs_val", "canonical", "none"])
def test_handle_lap_baselines(method):
    """The non-learned PE variants against oracle/restate.handle_lap (itself bit-exact against the reference)."""
    from signnet_basisnet_b200.gatedgcn_net import handle_lap

    d = synth_batch(11, "zinc", seed=6, k_dgl=8)
    model = types.SimpleNamespace(lap_method=method)
    gen = torch.Generator().manual_seed(5)
    out = handle_lap(model, d.pos_enc.to(DEV), _G(d.to(DEV)), DEV, generator=gen)
    flip = torch.rand(d.pos_enc.size(1), generator=torch.Generator().manual_seed(5))
    flip = torch.where(flip >= 0.5, 1.0, -1.0)
    ref = restate.handle_lap(d.pos_enc.clone(), d.num_nodes_per_graph, method, sign_flip=flip)
    assert torch.equal(out.cpu(), ref)

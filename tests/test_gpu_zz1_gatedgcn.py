"""GatedGCN predictor + PE baselines on the GPU against the reference's own output / gradients
(tests/golden/dgl_gatedgcn_net.pt) and the CPU oracle.

The CUDA side (csrc/gated.cu, signnet_basisnet_b200/gatedgcn_net.py) was written after the round's GPU budget was spent
and has not run on a GPU yet, so these tests run last in the GPU session as non-strict xfail (XPASS = parity observed); the oracle side is pinned by
tests/test_oracle_vs_reference.py / tests/test_oracle_golden.py on the CPU."""
import os
import types

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
    pe = handle_lap(net, d.pos_enc, G, DEV)                                # 'sign_inv', train_ZINC_graph_regression.py:20-25
    out, g_ret = net(G, d.x[:, 0], pe, d.edge_attr.reshape(-1), None)
    assert g_ret is G and out.shape == g["out"].shape
    assert_close_rel(out.detach().cpu(), g["out"], 2e-5, what="GatedGCNNet vs reference")
    (out * g["w"].to(DEV)).sum().backward()
    got = {k: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"])
    assert_grads_close(got, g["grads"], 1e-4, "GatedGCNNet vs reference")
    after = net.state_dict()
    for k, v in g["state_dict_after"].items():
        if "running_" in k and k.startswith("layers."):
            torch.testing.assert_close(after[k].cpu(), v, rtol=1e-5, atol=1e-6)


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

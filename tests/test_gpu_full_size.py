"""BASELINE.json's full size (cfg 4: 1024 ZINC-shape graphs, n_hid = 128, 8 phi layers, k = N_max = 37) through
size-independent properties - the CPU oracle would need tens of GB and minutes here:
  * aggregate (K1): linearity, and the column checksum  sum_i out_i = sum_j (1 + eps + outdeg_j) x_j  ("a checksum of
    checksums": every edge contributes its source row exactly once);
  * SignNetGNN (eval): exact sign invariance, batch-composition invariance (graphs are independent), finiteness;
  * one training step: gradients finite, BatchNorm running statistics updated twice per phi norm (+v, -v passes).
"""
import pytest
import torch

from signnet_basisnet_b200.synth import synth_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"
B_FULL, HID = 1024, 128


@pytest.fixture(scope="module")
def batch():
    return synth_batch(B_FULL, "zinc", seed=77)


def test_aggregate_linearity_and_checksum_full_size(batch):
    from signnet_basisnet_b200.layout import GraphIndex
    from signnet_basisnet_b200.phi import gin_agg

    d = batch
    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    sl = gi.slots_all(HID)
    assert not sl.use_generic_agg
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(2, sl.R, HID, device=DEV, generator=g)
    y = torch.randn(2, sl.R, HID, device=DEV, generator=g)
    eps = torch.tensor([0.25], device=DEV)
    ax, ay, axy = (torch.empty_like(x) for _ in range(3))
    gin_agg(x, ax, sl, 2, HID, eps=eps)
    gin_agg(y, ay, sl, 2, HID, eps=eps)
    gin_agg(2.0 * x - 0.5 * y, axy, sl, 2, HID, eps=eps)
    ref = 2.0 * ax - 0.5 * ay
    assert (axy - ref).abs().max() <= 1e-5 * ref.abs().max()
    # checksum: per (sign, graph, slot) block, column sums of the output equal the (1 + eps + outdeg)-weighted column
    # sums of the input; summed over everything and compared in fp64
    outdeg = torch.zeros(gi.N, dtype=torch.float64, device=DEV).index_add_(
        0, gi.edge_index[0], torch.ones(gi.E, dtype=torch.float64, device=DEV))
    # node id of every slot row: row(b, j, i) -> graph_ptr[b] + i
    rows = torch.arange(sl.R, device=DEV)
    b_of = torch.searchsorted(sl.row_ptr[1:].contiguous(), rows, right=True)
    n_b = (gi.graph_ptr[1:] - gi.graph_ptr[:-1]).to(torch.int64)
    local = (rows - sl.row_ptr[b_of]) % n_b[b_of]
    node = gi.graph_ptr[b_of].to(torch.int64) + local
    wgt = (1.0 + 0.25 + outdeg[node]).unsqueeze(-1)
    for s in (0, 1):
        lhs = ax[s].double().sum(0)
        rhs = (x[s].double() * wgt).sum(0)
        assert (lhs - rhs).abs().max() <= 1e-6 * rhs.abs().max().clamp(min=1.0) + 1e-3


def test_signnetgnn_full_size_invariances(batch):
    from signnet_basisnet_b200.sign_net import SignNetGNN

    d = batch
    torch.manual_seed(0)
    model = SignNetGNN(None, None, HID, 1, 8, 6, flavour="zinc").to(DEV)
    with torch.no_grad():
        for n_, b in model.named_buffers():
            if n_.endswith("running_mean"):
                b.normal_(0, 0.1)
            elif n_.endswith("running_var"):
                b.uniform_(0.5, 1.5)
    model.eval()
    dd = d.to(DEV)
    with torch.no_grad():
        out = model(dd)
        assert out.shape == (B_FULL, 1) and torch.isfinite(out).all()
        flipped = d.to(DEV)
        flipped.eigen_vectors = -flipped.eigen_vectors
        assert torch.equal(model(flipped), out), "sign invariance must be exact in eval mode"
        # the first 100 graphs alone (a different batch composition, different tiling) give the same rows
        nb = d.num_nodes_per_graph
        n100, v100 = int(nb[:100].sum()), int((nb[:100] ** 2).sum())
        e100 = int((d.edge_index[0] < n100).sum())
        sub = type(d)(x=d.x[:n100], edge_index=d.edge_index[:, :e100], edge_attr=d.edge_attr[:e100], batch=d.batch[:n100],
                      eigen_values=d.eigen_values[:n100], eigen_vectors=d.eigen_vectors[:v100], num_graphs=100)
        out100 = model(sub.to(DEV))
        assert (out100 - out[:100]).abs().max() <= 2e-5 * out.abs().max()


def test_training_step_full_size(batch):
    from signnet_basisnet_b200.sign_net import SignNetGNN

    torch.manual_seed(0)
    model = SignNetGNN(None, None, HID, 1, 8, 6, flavour="zinc").to(DEV).train()
    dd = batch.to(DEV)
    out = model(dd)
    out.abs().mean().backward()
    bad = [n for n, p in model.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    assert not bad, bad
    got = sum(p.grad is not None for p in model.parameters())
    assert got > 100
    # phi's BatchNorms see the +v pass and then the -v pass: two running-statistics updates per step
    assert int(model.sign_net.phi.norms[0].bn.num_batches_tracked) == 2
    assert int(model.gnn.norms[0].num_batches_tracked) == 1

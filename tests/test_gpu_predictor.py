"""Parity of the GINE predictor (a12) and its kernels against the CPU oracle."""
import pytest
import torch

import restate
from helpers import assert_close_rel, assert_grads_parity, assert_parity
from signnet_basisnet_b200.synth import synth_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _cpu_sd(module, dtype=torch.float32):
    sd = {k: (v.detach().cpu().clone().to(dtype) if v.is_floating_point() else v.detach().cpu().clone())
          for k, v in module.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    return sd


def _grads(sd):
    # shared parameters appear under two keys (convs.l.nn.* and convs.l.layer.nn.*); autograd fills the one used
    return {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}


@pytest.mark.parametrize("C", [128, 64, 20])
def test_gine_aggregate_bit_exact(C):
    from signnet_basisnet_b200.layout import GraphIndex, pad4
    from signnet_basisnet_b200.model import GineAggFn

    d = synth_batch(30, "zinc", seed=21)
    perm = torch.randperm(d.edge_index.shape[1], generator=torch.Generator().manual_seed(3))
    d.edge_index = d.edge_index[:, perm]
    N, E = d.batch.numel(), d.edge_index.shape[1]
    g = torch.Generator().manual_seed(4)
    x, e, eps = torch.randn(N, C, generator=g), torch.randn(E, C, generator=g), torch.tensor([0.21])
    ref = restate.gine_aggregate(x, d.edge_index, e, eps)
    ld = pad4(C)
    xp, ep = torch.zeros(N, ld), torch.zeros(E, ld)
    xp[:, :C], ep[:, :C] = x, e
    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    out = GineAggFn.apply(xp.to(DEV), ep.to(DEV), eps.to(DEV), gi).cpu()
    assert torch.equal(out[:, :C], ref)
    assert out[:, C:].abs().max() == 0 if ld > C else True


def test_gine_aggregate_backward():
    from signnet_basisnet_b200.layout import GraphIndex
    from signnet_basisnet_b200.model import GineAggFn

    d = synth_batch(12, "zinc", seed=22)
    N, E, C = d.batch.numel(), d.edge_index.shape[1], 32
    g = torch.Generator().manual_seed(5)
    x, e = torch.randn(N, C, generator=g), torch.randn(E, C, generator=g)
    eps, w = torch.tensor([-0.1]), torch.randn(N, C, generator=g)
    xr, er, epr = (t.double().requires_grad_(True) for t in (x, e, eps))
    (restate.gine_aggregate(xr, d.edge_index, er, epr) * w.double()).sum().backward()
    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    xg, eg, epg = (t.to(DEV).requires_grad_(True) for t in (x, e, eps))
    (GineAggFn.apply(xg, eg, epg, gi) * w.to(DEV)).sum().backward()
    assert_close_rel(xg.grad.cpu(), xr.grad.float(), TOL, what="dx")
    assert_close_rel(eg.grad.cpu(), er.grad.float(), TOL, what="de")
    assert_close_rel(epg.grad.cpu(), epr.grad.float(), TOL, what="deps")


@pytest.mark.parametrize("shape,B,nhid,nl", [("alchemy", 6, 16, 2), ("zinc", 5, 24, 3)])
def test_gnn_predictor(shape, B, nhid, nl):
    from signnet_basisnet_b200.model import GNN

    torch.manual_seed(5)
    d = synth_batch(B, shape, seed=23)
    nf, ef = (6, 4) if shape == "alchemy" else (None, None)
    if shape == "zinc":
        d.x, d.edge_attr = d.x % 6, d.edge_attr % 6
    net = GNN(nf, ef, nhid, 3, nl, "GINEConv").to(DEV).train()
    pos = torch.randn(d.batch.numel(), nhid, generator=torch.Generator().manual_seed(6))
    sd, sd64 = _cpu_sd(net), _cpu_sd(net, torch.float64)
    ref = restate.gnn_predictor(d.x, d.edge_index, d.edge_attr, d.batch, pos, sd, "", nl, d.num_graphs)
    ref.abs().sum().backward()
    d64 = d.edge_attr.double() if d.edge_attr.is_floating_point() else d.edge_attr
    x64 = d.x.double() if d.x.is_floating_point() else d.x
    ref64 = restate.gnn_predictor(x64, d.edge_index, d64, d.batch, pos.double(), sd64, "", nl, d.num_graphs)
    ref64.abs().sum().backward()
    dd = d.to(DEV)
    out = net(dd, pos.to(DEV))
    assert out.shape == ref.shape
    assert_parity(out, ref, ref64, TOL, what="GNN predictor")
    out.abs().sum().backward()
    got = {n_: p.grad.cpu() for n_, p in net.named_parameters() if p.grad is not None}
    r32 = {k: v for k, v in _grads(sd).items() if ".layer.nn." not in k}
    r64 = {k: v for k, v in _grads(sd64).items() if ".layer.nn." not in k}
    assert_grads_parity(got, r32, r64, TOL, "GNN predictor")
    for name, buf in net.named_buffers():
        if buf.is_floating_point():
            assert_parity(buf, sd[name], sd64[name], TOL, what=f"buffer {name}")


def test_state_dict_keys_match_reference_listing():
    """SURVEY §8b key listing (PyG flavour)."""
    from signnet_basisnet_b200.model import GNN

    keys = set(GNN(6, 4, 8, 3, 2).state_dict().keys())
    for k in ("input_encoder.layers.0.weight", "edge_encoders.1.norms.0.running_var", "convs.0.nn.layers.1.weight",
              "convs.1.nn.norms.1.num_batches_tracked", "convs.0.layer.eps", "convs.0.layer.nn.layers.0.weight",
              "norms.1.bias", "linear.bias", "output_encoder.layers.1.bias"):
        assert k in keys, k


@pytest.mark.parametrize("shape,nf,ef,nhid,L,training", [("alchemy", 6, 4, 64, 5, True), ("zinc", None, None, 95, 3, True),
                                                         ("alchemy", 6, 4, 20, 2, False), ("zinc", None, None, 128, 4, True)])
def test_gine_stack_driver_matches_per_module_path(shape, nf, ef, nhid, L, training, monkeypatch):
    """sb_gine_stack_fwd / _bwd (the predictor's layer loop in host C++, csrc/gine_stack.cu) issues the same kernels in the
    same order as the per-module Python path (SB_GINE_PER_CALL=1): outputs, every gradient and every BatchNorm buffer must
    be bit-identical; and far fewer C-ABI calls."""
    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200.model import GNN

    torch.manual_seed(11)
    d = synth_batch(17, shape, seed=41).to(DEV)
    if shape == "zinc":
        d.x, d.edge_attr = d.x % 6, d.edge_attr % 6
    ref_net = GNN(nf, ef, nhid, 3, L).to(DEV)
    pos = torch.randn(d.batch.numel(), nhid, device=DEV)
    w = torch.randn(17, 3, device=DEV)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SB_GINE_PER_CALL", mode)
        net = GNN(nf, ef, nhid, 3, L).to(DEV)
        net.load_state_dict(ref_net.state_dict())
        net.train(training)
        p = pos.clone().requires_grad_(True)
        c0 = _lib.launch_count
        out = net(d, p)
        (out * w).sum().backward()
        torch.cuda.synchronize()
        res[mode] = (out.detach().clone(), p.grad.clone(), {k: v.grad.clone() for k, v in net.named_parameters() if v.grad is not None},
                     {k: v.clone() for k, v in net.named_buffers()}, _lib.launch_count - c0)
    a, b = res["1"], res["0"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert set(a[2]) == set(b[2])
    for k in a[2]:
        assert torch.equal(a[2][k], b[2][k]), k
    for k in a[3]:
        assert torch.equal(a[3][k], b[3][k]), k
    assert b[4] < a[4], (a[4], b[4])   # e.g. 618 -> ~150 per step for the 16-layer Alchemy predictor

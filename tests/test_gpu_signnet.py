"""Parity of the CUDA SignNet path (modules -> autograd Functions -> C ABI) against the CPU oracle on seeded inputs:
outputs and every parameter gradient within 1e-5 relative fp32 (BASELINE.json north_star), BatchNorm running buffers
included."""
import pytest
import torch

import restate
from helpers import (assert_close_rel, assert_grads_parity, assert_parity, dense_to_rows, rows_to_dense,
                     slot_row_index)
from signnet_basisnet_b200.synth import synth_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _cpu_sd(module, leaf=True, dtype=torch.float32):
    sd = {k: (v.detach().cpu().clone().to(dtype) if v.is_floating_point() else v.detach().cpu().clone())
          for k, v in module.state_dict().items()}
    if leaf:
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
    return sd


def _check_buffers(module, sd, sd64, what):
    for name, buf in module.named_buffers():
        if buf.is_floating_point():
            assert_parity(buf, sd[name], sd64[name], TOL, what=f"{what} buffer {name}")
        else:
            assert torch.equal(buf.cpu(), sd[name]), f"{what} buffer {name}: {buf} vs {sd[name]}"


def _grads(sd):
    return {n_: v.grad for n_, v in sd.items() if v.requires_grad and v.grad is not None}


@pytest.mark.parametrize("shape,B,flavour,nhid,nl", [("alchemy", 24, "alchemy", 64, 3), ("zinc", 16, "alchemy", 128, 2),
                                                     ("alchemy", 128, "alchemy", 64, 8), ("zinc", 64, "zinc", 128, 8),
                                                     ("zinc", 12, "zinc", 95, 3), ("alchemy", 9, "zinc", 20, 4)])
def test_phi_stack_forward_backward(shape, B, flavour, nhid, nl):
    from signnet_basisnet_b200.layout import GraphIndex, pad4
    from signnet_basisnet_b200.sign_net import GNN3d, build_phi_input

    torch.manual_seed(0)
    d = synth_batch(B, shape, seed=11)
    phi = GNN3d(1, nhid, nl, flavour=flavour).to(DEV).train()
    with torch.no_grad():  # non-trivial affine parameters / eps so every gradient path is exercised
        for n_, p in phi.named_parameters():
            if n_.endswith("bn.weight"):
                p.uniform_(0.5, 1.5)
            elif n_.endswith("bn.bias") or n_.endswith("eps"):
                p.uniform_(-0.3, 0.3)
    _, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
    k = eigV.shape[1]
    mask = restate.slot_mask(d.batch, k)
    w = torch.randn(eigV.shape[0], k, nhid, generator=torch.Generator().manual_seed(1)) * mask.unsqueeze(-1)
    sd, sd64 = _cpu_sd(phi), _cpu_sd(phi, dtype=torch.float64)

    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    sl = gi.slots_all(pad4(nhid))
    assert sl.k == k
    x0 = build_phi_input(gi, sl, d.eigen_vectors.to(DEV))
    cap = []
    xr, sl = phi.forward_rows(x0, gi, k, True, capture=cap)
    idx = slot_row_index(d.batch, k, True)

    # Activation patterns of the CUDA pass, imposed on the oracle (see restate.ACTIVATION_OVERRIDE): pattern[tag] is
    # the [k, N, C] 0/1 mask of one ReLU.  They must agree with the oracle's own pattern except where the exact (fp64)
    # pre-activation is within rounding distance of zero.
    pattern, n_flip = {}, [0]
    for l, c_ in enumerate(cap):
        for key, a_, c0_, C, tag in ((c_["H"], c_["a0"], c_["c0"], c_["h"], f"convs.{l}.nn.norms.0"),
                                     (c_["Y"], c_["a1"], c_["c1"], c_["d"], f"norms.{l}")):
            z = torch.addcmul(c0_[:, None, :], a_[:, None, :], key[..., :C])
            for s_, sign in enumerate("+-"):
                pattern[sign + tag] = rows_to_dense((z[s_] > 0).float().cpu(), idx, C).transpose(0, 1)

    def act(dtype):
        def f(z, tag):
            m = pattern[tag].to(dtype)
            flips = ((z > 0).to(dtype) != m) & mask.transpose(0, 1).unsqueeze(-1)
            if flips.any():
                n_flip[0] += int(flips.sum())
                assert float(z.detach()[flips].abs().max()) <= 1e-5 * float(z.detach().abs().max()), f"activation pattern differs at {tag}"
            return z * m
        return f

    try:
        restate.ACTIVATION_OVERRIDE = act(torch.float32)
        ref = restate.phi_pm(eigV, d.edge_index, mask, sd, "", nl, True)
        (ref * w).sum().backward()
        restate.ACTIVATION_OVERRIDE = act(torch.float64)
        ref64 = restate.phi_pm(eigV.double(), d.edge_index, mask, sd64, "", nl, True)
        (ref64 * w.double()).sum().backward()
    finally:
        restate.ACTIVATION_OVERRIDE = None
    got = rows_to_dense(xr[0].cpu(), idx, nhid) + rows_to_dense(xr[1].cpu(), idx, nhid)
    assert_parity(got, ref, ref64, TOL, what="phi(+v)+phi(-v)")
    w_rows = dense_to_rows(w, idx, pad4(nhid)).to(DEV)
    (xr * w_rows.unsqueeze(0)).sum().backward()
    assert_grads_parity({n_: p.grad.cpu() for n_, p in phi.named_parameters() if p.grad is not None},
                        _grads(sd), _grads(sd64), TOL, "phi")
    _check_buffers(phi, sd, sd64, "phi")


def test_phi_eval_mode_and_sign_invariance():
    from signnet_basisnet_b200.layout import GraphIndex, pad4
    from signnet_basisnet_b200.sign_net import GNN3d, build_phi_input

    torch.manual_seed(3)
    d = synth_batch(10, "zinc", seed=12)
    phi = GNN3d(1, 32, 2).to(DEV)
    with torch.no_grad():
        for n_, b in phi.named_buffers():
            if n_.endswith("running_mean"):
                b.normal_(0, 0.2)
            elif n_.endswith("running_var"):
                b.uniform_(0.5, 2.0)
    phi.eval()
    sd = _cpu_sd(phi, leaf=False)
    _, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
    k = eigV.shape[1]
    mask = restate.slot_mask(d.batch, k)
    ref = restate.phi_pm(eigV, d.edge_index, mask, sd, "", 2, False)
    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    sl = gi.slots_all(pad4(32))
    idx = slot_row_index(d.batch, k, True)
    outs = []
    for sign in (1.0, -1.0):
        x0 = build_phi_input(gi, sl, (sign * d.eigen_vectors).to(DEV))
        with torch.no_grad():
            xr, _ = phi.forward_rows(x0, gi, k, True)
        outs.append(rows_to_dense(xr[0].cpu(), idx, 32) + rows_to_dense(xr[1].cpu(), idx, 32))
    assert_close_rel(outs[0], ref, TOL, what="phi eval")
    # exact sign invariance in eval mode (SURVEY §4 [probe]: max|f(V) - f(-V)| = 0)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("shape,B,ignore_eigval", [("alchemy", 20, False), ("zinc", 10, True)])
def test_signnet_rho0_module(shape, B, ignore_eigval):
    """SignNet with nl_rho = 0 (rho = sum over slots -> Linear -> BN) end to end through forward(data)."""
    from signnet_basisnet_b200.sign_net import SignNet

    torch.manual_seed(4)
    d = synth_batch(B, shape, seed=13)
    net = SignNet(48, 3, nl_rho=0, ignore_eigval=ignore_eigval).to(DEV).train()
    sd, sd64 = _cpu_sd(net), _cpu_sd(net, dtype=torch.float64)
    eigS, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
    ref = restate.sign_net(eigS, eigV, d.edge_index, d.batch, sd, "", 3, 0, ignore_eigval, True)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    (ref * w).sum().backward()
    ref64 = restate.sign_net(eigS.double(), eigV.double(), d.edge_index, d.batch, sd64, "", 3, 0, ignore_eigval, True)
    (ref64 * w.double()).sum().backward()
    out = net(d.to(DEV))
    assert out.shape == ref.shape
    assert_parity(out, ref, ref64, TOL, what="SignNet(nl_rho=0)")
    (out * w.to(DEV)).sum().backward()
    assert_grads_parity({n_: p.grad.cpu() for n_, p in net.named_parameters() if p.grad is not None},
                        _grads(sd), _grads(sd64), TOL, "signnet")
    _check_buffers(net, sd, sd64, "signnet")
    # tensor-level overload: forward(x, edge_index, eigvecs[N,k], batch, edge_attr, eigvals[N,k])
    net2 = SignNet(48, 3, nl_rho=0, ignore_eigval=ignore_eigval).to(DEV).train()
    net2.load_state_dict({k_: v.detach() for k_, v in _cpu_sd(net, False).items()})
    net.load_state_dict(net2.state_dict())
    dd = d.to(DEV)
    o1 = net(dd)
    o2 = net2(None, dd.edge_index, eigV.to(DEV), dd.batch, None, eigS.to(DEV))
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("masked", [True, False])
def test_gnn3d_forward_single_sign_pass(masked):
    """The reference's own call GNN3d.forward(x[N,k,1], edge_index, edge_attr, mask[N,k]) (sign_net.py:28-44): ONE sign
    pass, dense [N,k,d] result with zeros in the masked slots, one BatchNorm update - vs restate.gnn3d, fwd + grads."""
    from signnet_basisnet_b200.sign_net import GNN3d

    torch.manual_seed(5)
    d = synth_batch(14, "alchemy", seed=17)
    nhid, nl = 32, 3
    phi = GNN3d(1, nhid, nl).to(DEV).train()
    _, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
    k = eigV.shape[1] if masked else 4
    eigV = eigV[:, :k].contiguous()
    mask = restate.slot_mask(d.batch, k) if masked else None
    sd, sd64 = _cpu_sd(phi), _cpu_sd(phi, dtype=torch.float64)
    w = torch.randn(eigV.shape[0], k, nhid, generator=torch.Generator().manual_seed(6))
    if masked:
        w = w * mask.unsqueeze(-1)
    ref = restate.gnn3d(eigV.unsqueeze(-1), d.edge_index, mask, sd, "", nl, True)
    (ref * w).sum().backward()
    ref64 = restate.gnn3d(eigV.double().unsqueeze(-1), d.edge_index, mask, sd64, "", nl, True)
    (ref64 * w.double()).sum().backward()
    out = phi(eigV.unsqueeze(-1).to(DEV), d.edge_index.to(DEV), None, None if mask is None else mask.to(DEV),
              batch=d.batch.to(DEV))
    assert out.shape == ref.shape
    assert_parity(out, ref, ref64, TOL, what="GNN3d.forward")
    if masked:
        assert float(out[~mask.to(DEV)].abs().max()) == 0.0
    (out * w.to(DEV)).sum().backward()
    assert_grads_parity({n_: p.grad.cpu() for n_, p in phi.named_parameters() if p.grad is not None},
                        _grads(sd), _grads(sd64), TOL, "GNN3d.forward")
    _check_buffers(phi, sd, sd64, "GNN3d.forward")
    if masked:
        with pytest.raises(ValueError):
            phi(eigV.unsqueeze(-1).to(DEV), d.edge_index.to(DEV), None, torch.ones_like(mask).to(DEV), batch=d.batch.to(DEV))

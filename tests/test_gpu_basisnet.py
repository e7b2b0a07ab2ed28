"""BasisNet IGN phi (cfg 5) on the GPU: the 2->1 contractions from eigenvector factors / projectors and
the IGN2to1 module against the CPU oracle (oracle/restate.py) and the golden fixture written by the reference's own
IGN2to1 (tests/golden/ign2to1.pt, oracle/make_golden.py)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import restate  # noqa: E402
from helpers import assert_close_rel, assert_grads_parity, assert_parity  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _eigenspaces(n, mult, b, seed):
    g = torch.Generator().manual_seed(seed)
    V = torch.linalg.qr(torch.randn(n, mult * b + 3, generator=g))[0].contiguous()
    col0 = torch.arange(b, dtype=torch.int32) * mult + 1           # blocks start at column 1 (not 16-byte aligned)
    P = torch.stack([V[:, c:c + mult] @ V[:, c:c + mult].T for c in col0.tolist()]).unsqueeze(1)
    return V, col0, P


@pytest.mark.parametrize("n,mult,b", [(20, 2, 3), (333, 1, 5), (1000, 3, 4), (64, 32, 1)])
def test_ign_contractions_from_factors_and_projectors(n, mult, b):
    from signnet_basisnet_b200.basisnet import IGN2to1

    V, col0, P = _eigenspaces(n, mult, b, seed=n + mult)
    ref = restate.ign_2to1_ops(P.double())[:, 0].transpose(1, 2).reshape(b * n, 5)   # [b, 5, n] -> rows
    scale = float(ref.abs().max())
    for name, ops in (("factors", IGN2to1.ops_from_factors(V.to(DEV), col0.to(DEV), mult)),
                      ("projectors", IGN2to1.ops_from_projectors(P.to(DEV)))):
        ops = ops.cpu()
        assert ops.shape == (b * n, 8) and ops[:, 5:].abs().max() == 0
        # row/column sums of a projector orthogonal to the constant vector are pure rounding noise: absolute floor
        assert (ops[:, :5].double() - ref).abs().max() <= 2e-6 * scale, name


def test_ign2to1_golden(golden_dir):
    from signnet_basisnet_b200.basisnet import IGN2to1

    g = torch.load(os.path.join(golden_dir, "ign2to1.pt"))
    net = IGN2to1(1, 8, 2).to(DEV).train()
    missing = net.load_state_dict(g["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    out = net(g["P"].to(DEV)).cpu()
    assert_close_rel(out, g["out"], 1e-5, what="IGN2to1 (projectors) vs reference")
    net.load_state_dict(g["state_dict"])
    out_f = net.forward_factors(g["V"].to(DEV), torch.tensor([0, 2, 4], dtype=torch.int32, device=DEV), 2).cpu()
    assert_close_rel(out_f, g["out"], 1e-5, what="IGN2to1 (factors) vs reference")


@pytest.mark.parametrize("training", [True, False])
def test_ign2to1_forward_backward_vs_oracle(training):
    from signnet_basisnet_b200.basisnet import IGN2to1

    n, mult, b, hid = 150, 3, 6, 16
    V, col0, P = _eigenspaces(n, mult, b, seed=5)
    torch.manual_seed(1)
    net = IGN2to1(1, hid, mult).to(DEV)
    with torch.no_grad():   # non-trivial BN affine / running statistics, non-zero equivariant biases
        for bn in net.bns:
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)
            bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
        for lyr in net.equi_layers:
            lyr.bias.normal_(0, 0.1)
    net.train(training)
    def leaves(dtype):
        d = {k: (v.detach().cpu().to(dtype) if v.is_floating_point() else v.detach().cpu().clone())
             for k, v in net.state_dict().items()}
        for k, v in d.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
        return d

    sd, sd64 = leaves(torch.float32), leaves(torch.float64)
    w = torch.randn(b, mult, n, generator=torch.Generator().manual_seed(2))
    ref = restate.ign2to1(P, sd, training=training)
    (ref * w).sum().backward()
    P64 = torch.stack([V[:, c:c + mult].double() @ V[:, c:c + mult].double().T for c in col0.tolist()]).unsqueeze(1)
    ref64 = restate.ign2to1(P64, sd64, training=training)
    (ref64 * w.double()).sum().backward()
    out = net.forward_factors(V.to(DEV), col0.to(DEV), mult)
    (out * w.to(DEV)).sum().backward()
    # the repo's parity bar (tests/helpers.py): 1e-5 of the fp32 oracle, or as close to the exact (fp64) result as 4x
    # the fp32 oracle itself is - three BatchNorms after ReLUs amplify the rounding of the fp32 projector sums
    assert_parity(out, ref, ref64, 1e-5, what="IGN2to1 forward")
    got = {k: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
    g32 = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    g64 = {k: v.grad for k, v in sd64.items() if v.requires_grad and v.grad is not None}
    assert set(got) == set(g64)
    assert_grads_parity(got, g32, g64, 1e-5, "IGN2to1")
    if training:
        for i in range(3):
            assert_parity(net.bns[i].running_var, sd[f"bns.{i}.running_var"], sd64[f"bns.{i}.running_var"], 1e-5,
                          what="running_var")


def test_ign_basis_inv_and_grouping():
    from signnet_basisnet_b200.basisnet import IGNBasisInv, eigenspace_groups

    ev = torch.tensor([0.0, 0.5, 0.5000001, 1.0, 1.2, 1.2, 1.2, 2.0])
    groups = eigenspace_groups(ev)
    ref = restate.eigenspace_groups(ev)
    flat = sorted((int(s), m) for m, starts in groups.items() for s in starts.tolist())
    assert flat == sorted((a, b - a) for a, b in ref)
    net = IGNBasisInv(sorted(groups), 1, hidden_channels=8).to(DEV)
    V = torch.linalg.qr(torch.randn(40, 8, generator=torch.Generator().manual_seed(0)))[0].to(DEV)
    for m, starts in groups.items():
        y = net.forward_factors(V, starts.to(DEV), m)
        assert y.shape == (starts.numel(), m, 40) and torch.isfinite(y).all()
    with pytest.raises(ValueError):
        net.encs[0].ops_from_factors(V.cpu(), groups[1], 1)


@pytest.mark.parametrize("shape,cin,hid,cout,L", [((8, 200, 1), 1, 32, 1, 3), ((200, 16), 16, 10, 32, 3), ((5, 37, 3), 3, 12, 4, 1)])
def test_eq_deepsets_sign_plus_vs_oracle(shape, cin, hid, cout, L):
    """cfg 1 (row a14): SignPlus(EqDeepSetsEncoder) = phi of the single-graph SignNet on [k, n, 1]; rho on [n, 2k]."""
    from signnet_basisnet_b200.basisnet import EqDeepSetsEncoder, SignPlus

    torch.manual_seed(shape[-2])
    net = SignPlus(EqDeepSetsEncoder(cin, hid, cout, L, use_bn=True)).to(DEV).train()
    with torch.no_grad():
        for bn in getattr(net.model, "bns", []):
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.2)

    def leaves(dtype):
        d = {k[len("model."):]: v.detach().cpu().to(dtype) for k, v in net.state_dict().items()}
        for v in d.values():
            v.requires_grad_(True)
        return d

    sd, sd64 = leaves(torch.float32), leaves(torch.float64)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(*shape, generator=g)
    w = torch.randn(*shape[:-1], cout, generator=g)
    ref = restate.sign_plus_deepsets(x, sd, "", L)
    (ref * w).sum().backward()
    ref64 = restate.sign_plus_deepsets(x.double(), sd64, "", L)
    (ref64 * w.double()).sum().backward()
    out = net(x.to(DEV))
    (out * w.to(DEV)).sum().backward()
    assert_parity(out, ref, ref64, 1e-5, what="SignPlus(EqDeepSets)")
    got = {k[len("model."):]: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
    g32 = {k: v.grad for k, v in sd.items() if v.grad is not None}
    g64 = {k: v.grad for k, v in sd64.items() if v.grad is not None}
    assert set(got) == set(g64)
    assert_grads_parity(got, g32, g64, 1e-5, "SignPlus(EqDeepSets)")
    # sign invariance: f(v) == f(-v) up to the order of the fp64 statistics atomics
    a, b2 = net(x.to(DEV)), net(-x.to(DEV))
    assert (a - b2).abs().max() <= 1e-6 * a.abs().max()

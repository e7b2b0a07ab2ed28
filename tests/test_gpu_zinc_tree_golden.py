"""The GINESignNetPyG tree (`flavour='zinc'`: cfg 3 and the model bench.py times) on the GPU against the unmodified
reference's own output (tests/golden/zinc_pyg.pt, produced by GINESignNetPyG/core/sign_net.py SignNetGNN through
oracle/make_golden.py) and against gradients obtained by autograd through the CPU oracle of the same forward (the
reference's own backward does not run under torch 2.11: in-place writes on ReLU outputs, core/sign_net.py:46,
core/model.py:67)."""
import os

import pytest
import torch

import restate
from helpers import assert_grads_parity, assert_parity
from signnet_basisnet_b200.synth import Data

TOL = 1e-5

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _fatten(sd, rows_full=500):
    """DiscreteEncoder tables are stored with their first 32 rows only (make_golden._slim); the rest is never read."""
    out = {}
    for k, v in sd.items():
        if ".embeddings." in k and v.dim() == 2 and v.shape[0] < rows_full:
            v = torch.cat([v, torch.zeros(rows_full - v.shape[0], v.shape[1], dtype=v.dtype)])
        out[k] = v
    return out


def test_signnetgnn_zinc_tree_golden(golden_dir):
    from signnet_basisnet_b200.sign_net import SignNetGNN

    g = torch.load(os.path.join(golden_dir, "zinc_pyg.pt"), weights_only=False)
    d, c = Data(**g["data"]), g["cfg"]
    model = SignNetGNN(None, None, c["n_hid"], c["n_out"], c["nl_signnet"], c["nl_gnn"], flavour="zinc").to(DEV)
    sd = _fatten(g["state_dict"])
    assert set(model.state_dict()) == set(sd)
    model.load_state_dict(sd)   # a reference checkpoint loads unchanged
    for lyr in model.sign_net.rho.transformer_layers:
        lyr.slf_attn.attention.dropout.p = 0.0
    model.train()
    # fp64 arbiter: the oracle of the same forward, run in double precision on the same inputs
    sd64 = {k: (v.detach().clone().double() if v.is_floating_point() else v.detach().clone()) for k, v in sd.items()}
    for k, v in sd64.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    d64 = d.to("cpu")
    for k, v in list(d64.__dict__.items()):
        if torch.is_tensor(v) and v.is_floating_point():
            setattr(d64, k, v.double())
    ref64 = restate.sign_net_gnn(d64, sd64, c["nl_signnet"], c["nl_gnn"], nl_rho=1, ignore_eigval=True)
    (ref64 * g["w"].double()).sum().backward()
    out = model(d.to(DEV))
    assert_parity(out, g["out"], ref64, TOL, what="SignNetGNN (ZINC tree) vs reference")
    (out * g["w"].to(DEV)).sum().backward()
    want = _fatten(g["grads"])
    got = {k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None}
    missing = [k for k in want if k not in got and float(want[k].abs().max()) > 0]
    assert not missing, missing
    want = {k: v for k, v in want.items() if k in got}
    assert_grads_parity(got, want, {k: sd64[k].grad for k in want}, TOL, "SignNetGNN (ZINC tree) vs oracle autograd")
    after = model.state_dict()
    for k, v in g["buffers_after"].items():
        # eigen_encoder2 is evaluated and discarded by the reference (quirk v): only there do its statistics move
        if "running_" in k and "eigen_encoder" not in k and ".layer.nn." not in k:
            assert_parity(after[k], v, sd64[k], TOL, what=k)

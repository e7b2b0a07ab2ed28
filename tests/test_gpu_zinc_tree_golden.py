"""The GINESignNetPyG tree (`flavour='zinc'`: cfg 3 and the model bench.py times) on the GPU against the unmodified
reference's own output (tests/golden/zinc_pyg.pt, produced by GINESignNetPyG/core/sign_net.py SignNetGNN through
oracle/make_golden.py) and against gradients obtained by autograd through the CPU oracle of the same forward (the
reference's own backward does not run under torch 2.11: in-place writes on ReLU outputs, core/sign_net.py:46,
core/model.py:67).  Added after the round's last GPU visit; sorted last among the GPU test files on purpose."""
import os

import pytest
import torch

from helpers import assert_close_rel, assert_grads_close
from signnet_basisnet_b200.synth import Data

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


@pytest.mark.xfail(strict=False, reason="written after round 1's last GPU visit: outcome on a GPU not yet observed "
                                        "(XPASS = parity holds; remove this marker once seen)")
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
    out = model(d.to(DEV))
    assert_close_rel(out.cpu(), g["out"], 1e-5, what="SignNetGNN (ZINC tree) vs reference")
    (out * g["w"].to(DEV)).sum().backward()
    want = _fatten(g["grads"])
    got = {k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None}
    missing = [k for k in want if k not in got and float(want[k].abs().max()) > 0]
    assert not missing, missing
    assert_grads_close(got, {k: v for k, v in want.items() if k in got}, 5e-5, "SignNetGNN (ZINC tree) vs oracle autograd")
    after = model.state_dict()
    for k, v in g["buffers_after"].items():
        # eigen_encoder2 is evaluated and discarded by the reference (quirk v): only there do its statistics move
        if "running_" in k and "eigen_encoder" not in k and ".layer.nn." not in k:
            assert_close_rel(after[k].cpu(), v, 2e-5, what=k)

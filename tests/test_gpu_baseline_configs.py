"""Oracle parity of the full SignNetGNN forward + backward at the REAL depth and width of BASELINE.json's configs
(VERDICT r1 "next round" item 1).  The small-model tests elsewhere pin the arithmetic; these pin the error growth
through 8 BatchNorm'd phi layers under the 3xTF32 tensor-core contractions.

  cfg 2  Alchemy tree, B=128, n_hid=64, nl_signnet=8, nl_gnn=16, nl_rho=4, n_out=12   (main_alchemy.py:35,84)
  cfg 3  ZINC tree, B=256, n_hid=95, 4 phi layers, 6 GINE layers, k=8 eigenvector columns (unmasked: every ZINC-shape
         graph has n_b >= 9 > k), through the tensor overload forward(x, edge_index, eigvecs[N,k], batch, edge_attr)
         (core/config.py:57, zinc.yaml:9; k/hidden from configs/gin/GIN_ZINC_LapPE_signinv_GIN.json:27,36)
  cfg 4  ZINC tree at the benchmark's depth/width: n_hid=128, 8 phi layers, 6 GINE layers, k=N_max masked, B=64
         (the CPU oracle needs ~7 s for fp32 + fp64 at this size; B=1024 is covered by tests/test_gpu_full_size.py)

Bar: 1e-5 relative (BASELINE.json north_star) for the outputs AND every parameter gradient, with the fp64 run of the same
oracle as arbiter (helpers.assert_parity; which branch decided is printed at the end of the session).  At this depth the
fp32 ORACLE's own gradients are 1e-4..2e-3 away from its fp64 run (ReLU kinks within rounding distance, 8 BatchNorm
layers deep), so gradients use helpers.assert_grads_parity(family=True): the CUDA path must be no farther from exact
than the fp32 oracle is; the arithmetic proper is pinned at 1e-5 with activation patterns imposed, at the same depth
and width, by tests/test_gpu_signnet.py::test_phi_stack_forward_backward."""
import pytest
import torch

import restate
from helpers import assert_grads_parity, assert_parity, fp32_noise_samples
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
    return {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None and ".layer.nn." not in k}


def _f64(d):
    out = d.to("cpu")
    for k in ("x", "edge_attr", "eigen_values", "eigen_vectors"):
        v = getattr(out, k)
        if v.is_floating_point():
            setattr(out, k, v.double())
    return out


def _no_attn_dropout(model):
    for lyr in model.sign_net.rho.transformer_layers:  # reference quirk: attention dropout defaults to 0.1
        lyr.slf_attn.attention.dropout.p = 0.0


def _jitter_eigenvalues(d, seed=9):
    # eigenvalue 1 is a frequent multiple eigenvalue of tree-like graphs and sits exactly on the ReLU kink of
    # eigen_encoder's BN(1) (see tests/test_gpu_full_model.py): compare arithmetic, not that degeneracy
    d.eigen_values = d.eigen_values + 0.02 * torch.randn(d.eigen_values.shape, generator=torch.Generator().manual_seed(seed))


def _check(model, out, ref, ref64, sd, sd64, w, what, run32):
    assert out.shape == ref.shape
    assert_parity(out, ref, ref64, TOL, what=what)
    (out * w.to(out.device)).sum().backward()
    got = {n_: p.grad.cpu() for n_, p in model.named_parameters() if p.grad is not None}

    def run(sd_):
        (run32(sd_) * w).sum().backward()
        return _grads(sd_)

    assert_grads_parity(got, _grads(sd), _grads(sd64), TOL, what, family=True, samples=fp32_noise_samples(run, sd, 3))
    for name, buf in model.named_buffers():
        if buf.is_floating_point() and "eigen_encoder2" not in name:
            assert_parity(buf, sd[name], sd64[name], TOL, what=f"{what} buffer {name}")


def test_cfg2_alchemy_b128_h64_l8_g16():
    from signnet_basisnet_b200.sign_net import SignNetGNN

    torch.manual_seed(2)
    d = synth_batch(128, "alchemy", seed=202)
    _jitter_eigenvalues(d)
    model = SignNetGNN(6, 4, n_hid=64, n_out=12, nl_signnet=8, nl_gnn=16).to(DEV).train()
    _no_attn_dropout(model)
    sd, sd64 = _cpu_sd(model), _cpu_sd(model, torch.float64)
    w = torch.randn(128, 12, generator=torch.Generator().manual_seed(3))
    ref = restate.sign_net_gnn(d, sd, 8, 16)
    (ref * w).sum().backward()
    ref64 = restate.sign_net_gnn(_f64(d), sd64, 8, 16)
    (ref64 * w.double()).sum().backward()
    out = model(d.to(DEV))
    _check(model, out, ref, ref64, sd, sd64, w, "cfg2 SignNetGNN", lambda sd_: restate.sign_net_gnn(d, sd_, 8, 16))


def test_cfg3_zinc_b256_h95_k8():
    from signnet_basisnet_b200.sign_net import SignNetGNN

    torch.manual_seed(3)
    d = synth_batch(256, "zinc", seed=303)
    model = SignNetGNN(None, None, 95, 1, 4, 6, flavour="zinc").to(DEV).train()
    _no_attn_dropout(model)
    sd, sd64 = _cpu_sd(model), _cpu_sd(model, torch.float64)
    _, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
    eigV = eigV[:, :8].contiguous()
    assert int(d.num_nodes_per_graph.min()) > 8   # k = 8 is unmasked on ZINC-shape graphs
    w = torch.randn(256, 1, generator=torch.Generator().manual_seed(4))

    def oracle(sd_, dtype):
        pos = restate.sign_net(None, eigV.to(dtype), d.edge_index, d.batch, sd_, "sign_net.", 4, 1, True, True)
        return restate.gnn_predictor(d.x, d.edge_index, d.edge_attr, d.batch, pos, sd_, "gnn.", 6, num_graphs=256)

    ref = oracle(sd, torch.float32)
    (ref * w).sum().backward()
    ref64 = oracle(sd64, torch.float64)
    (ref64 * w.double()).sum().backward()
    dd = d.to(DEV)
    out = model(dd.x, dd.edge_index, eigV.to(DEV), dd.batch, dd.edge_attr)
    _check(model, out, ref, ref64, sd, sd64, w, "cfg3 SignNetGNN(k=8)", lambda sd_: oracle(sd_, torch.float32))


def test_cfg4_zinc_depth_width_b64():
    from signnet_basisnet_b200.sign_net import SignNetGNN

    torch.manual_seed(4)
    d = synth_batch(64, "zinc", seed=404)
    model = SignNetGNN(None, None, 128, 1, 8, 6, flavour="zinc").to(DEV).train()
    _no_attn_dropout(model)
    sd, sd64 = _cpu_sd(model), _cpu_sd(model, torch.float64)
    w = torch.randn(64, 1, generator=torch.Generator().manual_seed(5))
    ref = restate.sign_net_gnn(d, sd, 8, 6, nl_rho=1, ignore_eigval=True)
    (ref * w).sum().backward()
    ref64 = restate.sign_net_gnn(_f64(d), sd64, 8, 6, nl_rho=1, ignore_eigval=True)
    (ref64 * w.double()).sum().backward()
    out = model(d.to(DEV))
    _check(model, out, ref, ref64, sd, sd64, w, "cfg4 SignNetGNN(8x128)",
           lambda sd_: restate.sign_net_gnn(d, sd_, 8, 6, nl_rho=1, ignore_eigval=True))

"""CUDA path against the committed golden fixtures (outputs/gradients of the reference's own modules)."""
import os

import pytest
import torch

import restate
from helpers import (assert_close_rel, assert_grads_parity, assert_parity, rows_to_dense, slot_row_index)
from signnet_basisnet_b200.synth import Data

TOL = 1e-5   # BASELINE.json north_star; the reference's own fp32 output is the target, an fp64 run of the oracle the arbiter

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _leaf64(sd):
    """fp64 leaf copy of a fixture's state_dict: the 'exact' arbiter of helpers.assert_parity is the oracle run in fp64."""
    out = {k: (v.detach().clone().double() if v.is_floating_point() else v.detach().clone()) for k, v in sd.items()}
    for k, v in out.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    return out


def _g64(sd64, strip=""):
    return {k[len(strip):] if strip else k: v.grad for k, v in sd64.items() if v.requires_grad and v.grad is not None}


def _f64(d):
    out = d.to("cpu")
    for k, v in list(out.__dict__.items()):
        if torch.is_tensor(v) and v.is_floating_point():
            setattr(out, k, v.double())
    return out


def test_dense_list_evd_golden_bit_exact(golden_dir):
    from signnet_basisnet_b200 import ops

    g = _load(golden_dir, "alchemy_pyg.pt")
    d = Data(**g["data"]).to(DEV)
    S, V = ops.to_dense_list_EVD(d.eigen_values, d.eigen_vectors, d.batch)
    assert torch.equal(S.cpu(), g["dense_list_evd"]["eigS"]) and torch.equal(V.cpu(), g["dense_list_evd"]["eigV"])


def test_phi_golden(golden_dir):
    from signnet_basisnet_b200.layout import GraphIndex, pad4
    from signnet_basisnet_b200.sign_net import GNN3d, build_phi_input

    g = _load(golden_dir, "alchemy_pyg.pt")
    d, ph = Data(**g["data"]), g["phi"]
    nh, nl = ph["cfg"]["n_hid"], ph["cfg"]["n_layer"]
    phi = GNN3d(1, nh, nl).to(DEV).train()
    phi.load_state_dict(ph["state_dict"])
    dd = d.to(DEV)
    gi = GraphIndex(dd.edge_index, dd.batch, d.num_graphs)
    sl = gi.slots_all(pad4(nh))
    x0 = build_phi_input(gi, sl, dd.eigen_vectors)
    xr, sl = phi.forward_rows(x0, gi, sl.k, True)
    idx = slot_row_index(d.batch, sl.k, True)
    got = rows_to_dense(xr[0].cpu(), idx, nh) + rows_to_dense(xr[1].cpu(), idx, nh)
    assert_close_rel(got, ph["out"], 1e-5, what="phi vs reference")
    for k, v in ph["state_dict_after"].items():
        if "num_batches" in k:
            assert torch.equal(phi.state_dict()[k].cpu(), v), k


def test_signnetgnn_golden(golden_dir):
    from signnet_basisnet_b200.sign_net import SignNetGNN

    g = _load(golden_dir, "alchemy_pyg.pt")
    d, m = Data(**g["data"]), g["signnetgnn"]
    c = m["cfg"]
    model = SignNetGNN(c["node_feat"], c["edge_feat"], c["n_hid"], c["n_out"], c["nl_signnet"], c["nl_gnn"]).to(DEV)
    model.load_state_dict(m["state_dict"])  # a reference checkpoint loads unchanged
    for lyr in model.sign_net.rho.transformer_layers:
        lyr.slf_attn.attention.dropout.p = 0.0
    model.train()
    sd64 = _leaf64(m["state_dict"])
    ref64 = restate.sign_net_gnn(_f64(d), sd64, c["nl_signnet"], c["nl_gnn"])
    ref64.abs().sum().backward()
    out = model(d.to(DEV))
    assert_parity(out, m["out"], ref64, TOL, what="SignNetGNN vs reference (train)")
    out.abs().sum().backward()
    got = {k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None}
    g64 = {k: v for k, v in _g64(sd64).items() if k in m["grads"]}
    assert_grads_parity(got, m["grads"], g64, TOL, "SignNetGNN vs reference")
    model.eval()
    with torch.no_grad():
        out_e = model(d.to(DEV))
        sd_e64 = {k: (v.double() if v.is_floating_point() else v) for k, v in m["state_dict_after"].items()}
        ref_e64 = restate.sign_net_gnn(_f64(d), sd_e64, c["nl_signnet"], c["nl_gnn"], training=False)
    assert_parity(out_e, m["out_eval"], ref_e64, TOL, what="SignNetGNN vs reference (eval)")


class _G:
    def __init__(self, d):
        self.src, self.dst, self.n = d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph

    def edges(self):
        return self.src, self.dst

    def batch_num_nodes(self):
        return self.n


@pytest.mark.parametrize("name", ["gin", "masked_gin"])
def test_dgl_deepsigns_golden(golden_dir, name):
    from signnet_basisnet_b200.deepsigns import get_sign_inv_net

    g = _load(golden_dir, "dgl_deepsigns.pt")
    d, m, k = Data(**g["data"]).to(DEV), g[name], g["k"]
    net = get_sign_inv_net(dict(sign_inv_net=name, hidden_dim=m["cfg"]["hidden"], phi_out_dim=m["cfg"]["out"],
                                sign_inv_layers=m["cfg"]["layers"], pos_enc_dim=k, dropout=0.0,
                                sign_inv_activation="relu", device=DEV)).to(DEV).train()
    assert set(net.state_dict()) == set(m["state_dict"])
    net.load_state_dict(m["state_dict"])
    sd64, dc = _leaf64(m["state_dict"]), d.to("cpu")
    x64 = dc.pos_enc.unsqueeze(-1).double()
    if name == "gin":
        ref64 = restate.gin_deepsigns(x64, dc.edge_index[0], dc.edge_index[1], sd64, m["cfg"]["layers"], k)
    else:
        ref64 = restate.masked_gin_deepsigns(x64, dc.edge_index[0], dc.edge_index[1], dc.num_nodes_per_graph, sd64,
                                             m["cfg"]["layers"], k)
    (ref64 * m["w"].double()).sum().backward()
    out = net(_G(d), d.pos_enc.unsqueeze(-1))
    assert out.shape == m["out"].shape
    assert_parity(out, m["out"], ref64, TOL, what=f"{name} vs reference")
    (out * m["w"].to(DEV)).sum().backward()
    got = {k_: p.grad.cpu() for k_, p in net.named_parameters() if p.grad is not None}
    assert_grads_parity(got, m["grads"], {k_: v for k_, v in _g64(sd64).items() if k_ in m["grads"]}, TOL,
                        f"{name} vs reference")
    for k_, v in m["state_dict_after"].items():
        if "running_" in k_:
            assert_parity(net.state_dict()[k_], v, sd64[k_], TOL, what=k_)
        elif "num_batches" in k_:
            assert torch.equal(net.state_dict()[k_].cpu(), v), k_


def test_gin_net_golden(golden_dir):
    """Row a13: GINNet(net_params) with its sign_inv_net on the GPU vs the reference's own output and gradients."""
    from signnet_basisnet_b200.gin_net import GINNet

    g = _load(golden_dir, "dgl_gin_net.pt")
    d, prm = Data(**g["data"]).to(DEV), dict(g["params"], device=DEV)
    net = GINNet(prm).to(DEV).train()
    assert set(net.state_dict()) == set(g["state_dict"])
    net.load_state_dict(g["state_dict"])
    G = _G(d)
    sd64, dc = _leaf64(g["state_dict"]), d.to("cpu")
    sub = {k[len("sign_inv_net."):]: v for k, v in sd64.items() if k.startswith("sign_inv_net.")}
    pe64 = restate.masked_gin_deepsigns(dc.pos_enc.unsqueeze(-1).double(), dc.edge_index[0], dc.edge_index[1],
                                        dc.num_nodes_per_graph, sub, prm["sign_inv_layers"], prm["pos_enc_dim"]).squeeze(-1)
    ref64 = restate.gin_net(dc.x[:, 0], pe64, dc.edge_index[0], dc.edge_index[1], dc.num_nodes_per_graph, sd64, prm["L"],
                            prm["readout"])
    (ref64 * g["w"].double()).sum().backward()
    pe = net.sign_inv_net(G, d.pos_enc.unsqueeze(-1)).squeeze(-1)       # handle_lap, train_ZINC_graph_regression.py:20-25
    out, g_ret = net(G, d.x[:, 0], pe, torch.ones(d.edge_index.shape[1], 1, device=DEV), None)
    assert g_ret is G and out.shape == g["out"].shape
    assert_parity(out, g["out"], ref64, TOL, what="GINNet vs reference")
    (out * g["w"].to(DEV)).sum().backward()
    got = {k: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"])
    assert_grads_parity(got, g["grads"], {k: v for k, v in _g64(sd64).items() if k in g["grads"]}, TOL,
                        "GINNet vs reference")


@pytest.mark.parametrize("name", ["phi", "rho"])
def test_eq_deepsets_golden(golden_dir, name):
    """Row a14: SignPlus(EqDeepSetsEncoder) on the GPU vs the reference's own output and gradients."""
    from signnet_basisnet_b200.basisnet import EqDeepSetsEncoder, SignPlus

    m = _load(golden_dir, "eq_deepsets.pt")[name]
    c = m["cfg"]
    net = SignPlus(EqDeepSetsEncoder(c["cin"], c["hid"], c["cout"], c["L"], use_bn=True)).to(DEV).train()
    assert set(net.state_dict()) == set(m["state_dict"])
    net.load_state_dict(m["state_dict"])
    sd64 = _leaf64({k[len("model."):]: v for k, v in m["state_dict"].items()})
    ref64 = restate.sign_plus_deepsets(m["x"].double(), sd64, "", c["L"])
    (ref64 * m["w"].double()).sum().backward()
    out = net(m["x"].to(DEV))
    assert_parity(out, m["out"], ref64, TOL, what=f"SignPlus(EqDeepSets) {name} vs reference")
    (out * m["w"].to(DEV)).sum().backward()
    got = {k: p.grad.cpu() for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(m["grads"])
    assert_grads_parity(got, m["grads"], {"model." + k: v for k, v in _g64(sd64).items()}, TOL,
                        f"SignPlus(EqDeepSets) {name} vs reference")

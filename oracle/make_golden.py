"""Generates the committed golden fixtures under tests/golden/ by running the reference's OWN modules (unmodified,
imported from /root/reference through oracle/ref_loader.py with the stand-in L1 ops of oracle/shim/) on small seeded
inputs.  TEST INFRASTRUCTURE ONLY.  Run in the authoring container:   python oracle/make_golden.py
The fixtures travel to the GPU box (where /root/reference does not exist) and pin both the CPU oracle
(tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_golden.py)."""
import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from signnet_basisnet_b200.synth import synth_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _data_dict(d):
    return {k: v for k, v in d.__dict__.items() if torch.is_tensor(v) or isinstance(v, int)}


def _sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _grads(m):
    return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}


def make_alchemy():
    sn = ref_loader.alchemy()
    tr = ref_loader.alchemy_transform()
    torch.manual_seed(0)
    d = synth_batch(5, "alchemy", seed=41)
    d.eigen_values = d.eigen_values + 0.02 * torch.randn(d.eigen_values.shape, generator=torch.Generator().manual_seed(9))
    out = {"data": _data_dict(d)}
    S, V = tr.to_dense_list_EVD(d.eigen_values, d.eigen_vectors, d.batch)
    out["dense_list_evd"] = {"eigS": S, "eigV": V}
    # phi alone (GNN3d), +v and -v passes
    phi = sn.GNN3d(1, 16, 3)
    with torch.no_grad():
        for n_, p in phi.named_parameters():
            if n_.endswith("bn.weight"):
                p.uniform_(0.5, 1.5)
            elif n_.endswith("bn.bias") or n_.endswith("eps"):
                p.uniform_(-0.3, 0.3)
    sd0 = _sd(phi)
    mask = torch.arange(V.shape[1])[None, :] < torch.bincount(d.batch)[d.batch][:, None]
    x = V.unsqueeze(-1)
    y = phi(x, d.edge_index, None, mask) + phi(-x, d.edge_index, None, mask)
    w = torch.randn(y.shape, generator=torch.Generator().manual_seed(1)) * mask.unsqueeze(-1)
    (y * w).sum().backward()
    out["phi"] = {"state_dict": sd0, "out": y.detach(), "w": w, "grads": _grads(phi), "state_dict_after": _sd(phi),
                  "cfg": dict(n_hid=16, n_layer=3)}
    # full SignNetGNN (attention dropout set to 0: reference quirk, see oracle/restate.py)
    model = sn.SignNetGNN(6, 4, n_hid=16, n_out=3, nl_signnet=2, nl_gnn=2)
    for lyr in model.sign_net.rho.transformer_layers:
        lyr.slf_attn.attention.dropout.p = 0.0
    sd0 = _sd(model)
    o = model(copy.copy(d))
    o.abs().sum().backward()
    out["signnetgnn"] = {"state_dict": sd0, "out": o.detach(), "grads": _grads(model), "state_dict_after": _sd(model),
                         "cfg": dict(node_feat=6, edge_feat=4, n_hid=16, n_out=3, nl_signnet=2, nl_gnn=2)}
    model.eval()
    out["signnetgnn"]["out_eval"] = model(copy.copy(d)).detach()
    torch.save(out, os.path.join(OUT, "alchemy_pyg.pt"))


def make_dgl():
    import dgl

    ds, _, _ = ref_loader.graphprediction_layers()
    torch.manual_seed(1)
    k = 6
    d = synth_batch(4, "zinc", seed=42, k_dgl=k)
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    out = {"data": _data_dict(d), "k": k}
    for name, net in (("gin", ds.GINDeepSigns(1, 12, 4, 4, k, use_bn=True, dropout=0.0, activation="relu")),
                      ("masked_gin", ds.MaskedGINDeepSigns(1, 12, 12, 4, k, "cpu", use_bn=True, dropout=0.0,
                                                           activation="relu"))):
        sd0 = _sd(net)
        x = d.pos_enc.unsqueeze(-1)
        y = net(g, x)
        w = torch.randn(y.shape, generator=torch.Generator().manual_seed(2))
        (y * w).sum().backward()
        out[name] = {"state_dict": sd0, "out": y.detach(), "w": w, "grads": _grads(net), "state_dict_after": _sd(net),
                     "cfg": dict(hidden=12, out=(4 if name == "gin" else 12), layers=4)}
    torch.save(out, os.path.join(OUT, "dgl_deepsigns.pt"))


def make_gin_net():
    """Row a13: the reference's own GINNet (+ its masked_gin sign_inv_net) on a small seeded batch."""
    gn = ref_loader.gin_net()
    import dgl
    params = dict(num_atom_type=28, num_bond_type=4, hidden_dim=20, out_dim=20, in_feat_dropout=0.0, dropout=0.0, L=3,
                  readout="mean", batch_norm=True, residual=True, edge_feat=False, device="cpu", pe_init="lap_pe",
                  lap_method="sign_inv", lap_lspe=False, use_lapeig_loss=False, lambda_loss=0.0, alpha_loss=0.0,
                  pos_enc_dim=6, sign_inv_net="masked_gin", phi_out_dim=8, sign_inv_layers=3, sign_inv_activation="relu")
    torch.manual_seed(7)
    net = gn.GINNet(params)
    d = synth_batch(6, "zinc", seed=13, k_dgl=params["pos_enc_dim"])
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    sd0 = _sd(net)
    pe = net.sign_inv_net(g, d.pos_enc.unsqueeze(-1)).squeeze(-1)
    out, _ = net(g, d.x[:, 0], pe, torch.ones(d.edge_index.shape[1], 1), None)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (out * w).sum().backward()
    torch.save({"params": params, "state_dict": sd0, "data": _data_dict(d), "w": w, "out": out.detach(),
                "grads": _grads(net)}, os.path.join(OUT, "dgl_gin_net.pt"))


def make_gatedgcn_net():
    """the reference's own GatedGCNNet (+ its masked_gin sign_inv_net; the structure of
    configs/gatedgcn/GatedGCN_ZINC_LapPE_signinv_GIN_mask.json at a small width) on a small seeded batch."""
    gg = ref_loader.gatedgcn_net()
    import dgl
    params = dict(num_atom_type=28, num_bond_type=4, hidden_dim=20, out_dim=20, in_feat_dropout=0.0, dropout=0.0, L=3,
                  readout="mean", batch_norm=True, residual=True, edge_feat=True, device="cpu", pe_init="lap_pe",
                  lap_method="sign_inv", lap_lspe=False, use_lapeig_loss=False, lambda_loss=1.0, alpha_loss=1e-4,
                  pos_enc_dim=6, sign_inv_net="masked_gin", phi_out_dim=8, sign_inv_layers=3, sign_inv_activation="relu",
                  pe_aggregate="concat")
    torch.manual_seed(17)
    net = gg.GatedGCNNet(params)
    d = synth_batch(6, "zinc", seed=23, k_dgl=params["pos_enc_dim"])
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    sd0 = _sd(net)
    pe = net.sign_inv_net(g, d.pos_enc.unsqueeze(-1)).squeeze(-1)
    out, _ = net(g, d.x[:, 0], pe, d.edge_attr.reshape(-1), None)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (out * w).sum().backward()
    torch.save({"params": params, "state_dict": sd0, "data": _data_dict(d), "w": w, "out": out.detach(),
                "grads": _grads(net), "state_dict_after": _sd(net)}, os.path.join(OUT, "dgl_gatedgcn_net.pt"))


def make_pna_net():
    """the reference's own PNANet (+ masked_gin sign_inv_net; structure of
    configs/pna/PNA_ZINC_LapPE_signinv_GIN_mask.json at a small width) on a small seeded batch."""
    pn = ref_loader.pna_net()
    import dgl
    params = dict(num_atom_type=28, num_bond_type=4, hidden_dim=20, out_dim=20, in_feat_dropout=0.0, dropout=0.0, L=3,
                  readout="sum", graph_norm=True, batch_norm=True, residual=True, aggregators="mean max min std",
                  scalers="identity amplification attenuation", avg_d={"log": 1.1}, towers=5, divide_input_first=True,
                  divide_input_last=True, edge_feat=True, edge_dim=8, pretrans_layers=1, posttrans_layers=1, gru=False,
                  device="cpu", pe_init="lap_pe", lap_method="sign_inv", lap_lspe=False, use_lapeig_loss=False,
                  lambda_loss=1000, alpha_loss=1e-4, pos_enc_dim=6, sign_inv_net="masked_gin", phi_out_dim=8,
                  sign_inv_layers=3, sign_inv_activation="relu", pe_aggregate="concat")
    torch.manual_seed(23)
    net = pn.PNANet(params)
    d = synth_batch(6, "zinc", seed=27, k_dgl=params["pos_enc_dim"])
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    n = torch.as_tensor(d.num_nodes_per_graph)
    snorm_n = (1.0 / n.float().sqrt()).repeat_interleave(n).unsqueeze(1)
    sd0 = _sd(net)
    pe = net.sign_inv_net(g, d.pos_enc.unsqueeze(-1)).squeeze(-1)
    out, _ = net(g, d.x[:, 0], pe, d.edge_attr.reshape(-1), snorm_n)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (out * w).sum().backward()
    torch.save({"params": params, "state_dict": sd0, "data": _data_dict(d), "snorm_n": snorm_n, "w": w, "out": out.detach(),
                "grads": _grads(net), "state_dict_after": _sd(net)}, os.path.join(OUT, "dgl_pna_net.pt"))


def make_transformer_net():
    """the reference's own TransformerNet (+ gin sign_inv_net; structure of
    configs/transformer/Transformer_ZINC_LapPE_signinv_GIN.json at a small width) on a small seeded batch."""
    tn = ref_loader.transformer_net()
    import dgl
    params = dict(num_atom_type=28, num_bond_type=4, hidden_dim=16, out_dim=16, n_heads=4, full_graph=False,
                  in_feat_dropout=0.0, dropout=0.0, L=3, readout="sum", batch_norm=True, layer_norm=True, residual=True,
                  edge_feat=True, device="cpu", pe_init="lap_pe", lap_method="sign_inv", lap_lspe=False,
                  use_lapeig_loss=False, lambda_loss=1, alpha_loss=1e-4, pos_enc_dim=6, sign_inv_net="gin", phi_out_dim=4,
                  sign_inv_layers=3, sign_inv_activation="relu", pe_aggregate="concat")
    torch.manual_seed(29)
    net = tn.TransformerNet(params)
    d = synth_batch(6, "zinc", seed=31, k_dgl=params["pos_enc_dim"])
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    sd0 = _sd(net)
    pe = net.sign_inv_net(g, d.pos_enc.unsqueeze(-1)).squeeze(-1)
    out, _ = net(g, d.x[:, 0], pe, d.edge_attr.reshape(-1), None)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (out * w).sum().backward()
    torch.save({"params": params, "state_dict": sd0, "data": _data_dict(d), "w": w, "out": out.detach(),
                "grads": _grads(net), "state_dict_after": _sd(net)}, os.path.join(OUT, "dgl_transformer_net.pt"))


def _slim(sd, rows=32):
    """DiscreteEncoder tables have 500 rows per feature (core/model_utils/elements.py:22) of which the ZINC-shape inputs
    touch < 32: keep the fixture small by storing only the first `rows` rows (the tests zero-pad them back)."""
    return {k: (v[:rows].clone() if ".embeddings." in k and v.dim() == 2 and v.shape[0] > rows else v) for k, v in sd.items()}


def make_zinc_pyg():
    """The GINESignNetPyG tree (cfg 3; the model bench.py times).  `out` and `state_dict_after` come from the unmodified
    reference (core/sign_net.py SignNetGNN, training-mode forward under no_grad: its backward cannot run under torch 2.11
    because of in-place writes on ReLU outputs); `grads` are torch autograd through oracle/restate.sign_net_gnn, whose
    forward is checked here to reproduce the reference's output (asserted below)."""
    import restate
    sn = ref_loader.gine_signnet_pyg()
    torch.manual_seed(19)
    cfg = dict(n_hid=24, n_out=1, nl_signnet=3, nl_gnn=2)
    d = synth_batch(7, "zinc", seed=29)
    model = sn.SignNetGNN(None, None, **cfg).train()
    for lyr in model.sign_net.rho.transformer_layers:
        lyr.slf_attn.attention.dropout.p = 0.0
    with torch.no_grad():   # non-trivial BatchNorm affines / eps so that every parameter matters
        for n_, p in model.named_parameters():
            if n_.endswith("bn.weight"):
                p.uniform_(0.5, 1.5)
            elif n_.endswith("bn.bias") or n_.endswith("eps"):
                p.uniform_(-0.3, 0.3)
    sd0 = _sd(model)
    with torch.no_grad():
        out = model(copy.copy(d))
    sd = {k: v.clone() for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    out_o = restate.sign_net_gnn(d, sd, cfg["nl_signnet"], cfg["nl_gnn"], nl_rho=1, ignore_eigval=True)
    torch.testing.assert_close(out_o, out, rtol=1e-5, atol=1e-5)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(4))
    (out_o * w).sum().backward()
    grads = {k: v.grad.detach().clone() for k, v in sd.items() if v.requires_grad and v.grad is not None}
    after = {k: v for k, v in _sd(model).items() if "running_" in k or "num_batches" in k}
    torch.save({"cfg": cfg, "state_dict": _slim(sd0), "data": _data_dict(d), "w": w, "out": out.detach(),
                "grads": _slim(grads), "embedding_rows_kept": 32,
                "grads_source": "torch autograd through oracle/restate.py (reference backward not runnable)",
                "buffers_after": after}, os.path.join(OUT, "zinc_pyg.pt"))


def make_eq_deepsets():
    """Row a14: the reference's own SignPlus(EqDeepSetsEncoder) (phi on [k, n, 1] and rho on [n, 2k], training.py:207-218)."""
    models = ref_loader.learningfilters_models()
    _, sbn = ref_loader.learningfilters()
    out = {}
    for name, shape, (cin, hid, cout, L) in (("phi", (8, 40, 1), (1, 32, 1, 3)), ("rho", (40, 16), (16, 10, 32, 3))):
        torch.manual_seed(11)
        net = sbn.SignPlus(models.EqDeepSetsEncoder(cin, hid, cout, L, use_bn=True))
        x = torch.randn(*shape)
        y = net(x)
        w = torch.randn(y.shape, generator=torch.Generator().manual_seed(2))
        (y * w).sum().backward()
        out[name] = {"cfg": dict(cin=cin, hid=hid, cout=cout, L=L), "state_dict": _sd(net), "x": x, "w": w,
                     "out": y.detach(), "grads": _grads(net)}
    torch.save(out, os.path.join(OUT, "eq_deepsets.pt"))


def make_ign():
    ign, _ = ref_loader.learningfilters()
    torch.manual_seed(3)
    net = ign.IGN2to1(1, 8, 2, device="cpu")
    V = torch.linalg.qr(torch.randn(20, 6))[0]
    P = torch.stack([V[:, :2] @ V[:, :2].T, V[:, 2:4] @ V[:, 2:4].T, V[:, 4:6] @ V[:, 4:6].T]).unsqueeze(1)
    sd0 = _sd(net)
    y = net(P)
    torch.save({"state_dict": sd0, "V": V, "P": P, "out": y.detach()}, os.path.join(OUT, "ign2to1.pt"))


if __name__ == "__main__":
    assert ref_loader.available(), "needs the read-only reference mount"
    os.makedirs(OUT, exist_ok=True)
    make_alchemy()
    make_dgl()
    make_gin_net()
    make_gatedgcn_net()
    make_pna_net()
    make_transformer_net()
    make_zinc_pyg()
    make_eq_deepsets()
    make_ign()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))

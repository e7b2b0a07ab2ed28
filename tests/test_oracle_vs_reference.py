"""Pins oracle/restate.py against the reference's OWN modules (unmodified, imported from /root/reference with the
stand-in L1 ops).  Skipped where the read-only mount is absent (the GPU box); there the committed golden fixtures
(tests/test_oracle_golden.py) carry the pin."""
import copy

import pytest
import torch

import ref_loader
import restate
from helpers import assert_grads_close
from signnet_basisnet_b200.synth import synth_batch

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")


def _clone_sd(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def _leafify(sd):
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    return sd


def test_dense_list_evd_bit_exact():
    tr = ref_loader.alchemy_transform()
    for shape, B in (("alchemy", 17), ("zinc", 9)):
        d = synth_batch(B, shape, seed=3)
        S_ref, V_ref = tr.to_dense_list_EVD(d.eigen_values, d.eigen_vectors, d.batch)
        S, V = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
        assert torch.equal(S, S_ref) and torch.equal(V, V_ref)


@pytest.mark.parametrize("shape,B,nhid,nl", [("alchemy", 8, 16, 3), ("zinc", 5, 24, 2)])
def test_phi_pyg_forward_backward(shape, B, nhid, nl):
    sn = ref_loader.alchemy()
    torch.manual_seed(0)
    d = synth_batch(B, shape, seed=1)
    phi = sn.GNN3d(1, nhid, nl)
    sd = _leafify(_clone_sd(phi))
    _, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
    mask = restate.slot_mask(d.batch, eigV.shape[1])
    x = eigV.unsqueeze(-1)
    ref = phi(x, d.edge_index, None, mask) + phi(-x, d.edge_index, None, mask)
    out = restate.phi_pm(eigV, d.edge_index, mask, sd, "", nl, True)
    torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-6)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (out * w).sum().backward()
    assert_grads_close({k: v.grad for k, v in sd.items()}, {k: v.grad for k, v in phi.named_parameters()}, 1e-5, "phi")
    # running statistics (+v then -v) and counters
    for name, buf in phi.named_buffers():
        torch.testing.assert_close(sd[name], buf, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("shape,B", [("alchemy", 6), ("zinc", 4)])
def test_full_signnetgnn(shape, B):
    sn = ref_loader.alchemy()
    torch.manual_seed(1)
    d = synth_batch(B, shape, seed=2)
    nf, ef = (6, 4) if shape == "alchemy" else (None, None)
    if shape == "zinc":  # the Alchemy tree's DiscreteEncoder has 6 values per feature (elements.py:22)
        d.x, d.edge_attr = d.x % 6, d.edge_attr % 6
    model = sn.SignNetGNN(nf, ef, n_hid=16, n_out=3, nl_signnet=2, nl_gnn=3)
    for lyr in model.sign_net.rho.transformer_layers:  # reference quirk: attention dropout defaults to 0.1
        lyr.slf_attn.attention.dropout.p = 0.0
    sd = _leafify(_clone_sd(model))
    ref = model(copy.copy(d))
    out = restate.sign_net_gnn(d, sd, 2, 3)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    ref.abs().sum().backward()
    out.abs().sum().backward()
    assert_grads_close({k: v.grad for k, v in sd.items()}, {k: v.grad for k, v in model.named_parameters()}, 2e-5, "full")


@pytest.mark.parametrize("B,nhid,nl_signnet,nl_gnn", [(5, 16, 2, 3), (9, 24, 3, 2)])
def test_full_signnetgnn_zinc_tree(B, nhid, nl_signnet, nl_gnn):
    """The GINESignNetPyG tree (cfg 3; the model bench.py times): MaskedMLP hidden width = input width (Linear(1->1) in
    the first phi layer), no Linear biases, nl_rho = 1, eigenvalue encoder computed and discarded, DiscreteEncoder
    inputs (core/sign_net.py:12-134, core/model.py:9-79).  Forward + running statistics against the unmodified
    reference (its backward cannot run under torch 2.11: in-place writes on ReLU outputs)."""
    sn = ref_loader.gine_signnet_pyg()
    torch.manual_seed(3)
    d = synth_batch(B, "zinc", seed=5)
    model = sn.SignNetGNN(None, None, n_hid=nhid, n_out=1, nl_signnet=nl_signnet, nl_gnn=nl_gnn).train()
    for lyr in model.sign_net.rho.transformer_layers:  # reference quirk: attention dropout defaults to 0.1
        lyr.slf_attn.attention.dropout.p = 0.0
    sd = _clone_sd(model)
    with torch.no_grad():
        ref = model(copy.copy(d))
    out = restate.sign_net_gnn(d, sd, nl_signnet, nl_gnn, nl_rho=1, ignore_eigval=True)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    for name, buf in model.named_buffers():
        # eigen_encoder2 runs in the reference and its result is discarded (quirk v): its BN buffers move there only
        if buf.is_floating_point() and "eigen_encoder" not in name:
            torch.testing.assert_close(sd[name], buf, rtol=1e-5, atol=1e-6, msg=name)


@pytest.mark.parametrize("masked", [False, True])
def test_dgl_deepsigns(masked):
    import dgl

    ds, _, _ = ref_loader.graphprediction_layers()
    torch.manual_seed(2)
    k = 6
    d = synth_batch(5, "zinc", seed=4, k_dgl=k)
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    if masked:
        net = ds.MaskedGINDeepSigns(1, 12, 12, 4, k, "cpu", use_bn=True, dropout=0.0, activation="relu")
    else:
        net = ds.GINDeepSigns(1, 12, 4, 4, k, use_bn=True, dropout=0.0, activation="relu")
    sd = _leafify(_clone_sd(net))
    x = d.pos_enc.unsqueeze(-1)
    ref = net(g, x)
    if masked:
        out = restate.masked_gin_deepsigns(x, d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph, sd, 4, k)
    else:
        out = restate.gin_deepsigns(x, d.edge_index[0], d.edge_index[1], sd, 4, k)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-5)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (out * w).sum().backward()
    assert_grads_close({k: v.grad for k, v in sd.items()}, {k: v.grad for k, v in net.named_parameters()}, 5e-5, "dgl")


def test_ign2to1():
    ign, _ = ref_loader.learningfilters()
    torch.manual_seed(3)
    net = ign.IGN2to1(1, 8, 2, device="cpu")
    sd = _clone_sd(net)
    V = torch.linalg.qr(torch.randn(20, 6))[0]
    P = torch.stack([V[:, :2] @ V[:, :2].T, V[:, 2:4] @ V[:, 2:4].T, V[:, 4:6] @ V[:, 4:6].T]).unsqueeze(1)
    torch.testing.assert_close(restate.ign2to1(P, sd), net(P), rtol=1e-5, atol=1e-6)


GIN_NET_PARAMS = dict(num_atom_type=28, num_bond_type=4, hidden_dim=20, out_dim=20, in_feat_dropout=0.0, dropout=0.0, L=3,
                      readout="mean", batch_norm=True, residual=True, edge_feat=False, device="cpu", pe_init="lap_pe",
                      lap_method="sign_inv", lap_lspe=False, use_lapeig_loss=False, lambda_loss=0.0, alpha_loss=0.0,
                      pos_enc_dim=6, sign_inv_net="masked_gin", phi_out_dim=8, sign_inv_layers=3,
                      sign_inv_activation="relu")


@pytest.mark.parametrize("readout", ["mean", "sum"])
def test_gin_net_predictor(readout):
    """Row a13: the DGL GINNet predictor consuming the sign-invariant PE (gin_net.py:81-138, handle_lap :20-25)."""
    gn = ref_loader.gin_net()   # puts the stand-in dgl on sys.path
    import dgl

    torch.manual_seed(5)
    params = dict(GIN_NET_PARAMS, readout=readout)
    net = gn.GINNet(params)
    k = params["pos_enc_dim"]
    d = synth_batch(6, "zinc", seed=12, k_dgl=k)
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    sd = _leafify(_clone_sd(net))
    atoms = d.x[:, 0]
    pe = net.sign_inv_net(g, d.pos_enc.unsqueeze(-1)).squeeze(-1)          # handle_lap, train_ZINC_graph_regression.py:20-25
    ref, _ = net(g, atoms, pe, torch.ones(d.edge_index.shape[1], 1), None)
    pe_o = restate.masked_gin_deepsigns(d.pos_enc.unsqueeze(-1), d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph,
                                        {k2[len("sign_inv_net."):]: v for k2, v in sd.items() if k2.startswith("sign_inv_net.")},
                                        params["sign_inv_layers"], k).squeeze(-1)
    out = restate.gin_net(atoms, pe_o, d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph, sd, params["L"], readout)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-5)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (out * w).sum().backward()
    want = {k2: v.grad for k2, v in net.named_parameters() if v.grad is not None}
    got = {k2: v.grad for k2, v in sd.items() if v.requires_grad and v.grad is not None and not k2.endswith(".eps")}
    assert set(want) == set(got)   # dgl GINConv eps is a buffer; embedding_e never reaches the output (no gradient)
    assert_grads_close(got, want, 5e-5, "gin_net")


@pytest.mark.parametrize("shape,cin,hid,cout,L", [((8, 50, 1), 1, 32, 1, 3), ((50, 16), 16, 10, 32, 3), ((4, 9, 3), 3, 6, 2, 1)])
def test_eq_deepsets_sign_plus(shape, cin, hid, cout, L):
    """Row a14: SignPlus(EqDeepSetsEncoder) of the single-graph SignNet (signbasisnet.py:11-20, models.py:58-113)."""
    models = ref_loader.learningfilters_models()
    _, sbn = ref_loader.learningfilters()
    torch.manual_seed(4)
    net = sbn.SignPlus(models.EqDeepSetsEncoder(cin, hid, cout, L, use_bn=True))
    sd = _leafify({k[len("model."):]: v for k, v in _clone_sd(net).items()})
    x = torch.randn(*shape)
    ref = net(x)
    out = restate.sign_plus_deepsets(x, sd, "", L)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (out * w).sum().backward()
    want = {k[len("model."):]: v.grad for k, v in net.named_parameters() if v.grad is not None}
    got = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    assert set(want) == set(got)
    assert_grads_close(got, want, 2e-5, "eq_deepsets")


GATEDGCN_NET_PARAMS = dict(num_atom_type=28, num_bond_type=4, hidden_dim=20, out_dim=20, in_feat_dropout=0.0, dropout=0.0,
                           L=3, readout="mean", batch_norm=True, residual=True, edge_feat=True, device="cpu",
                           pe_init="lap_pe", lap_method="sign_inv", lap_lspe=False, use_lapeig_loss=False,
                           lambda_loss=1.0, alpha_loss=1e-4, pos_enc_dim=6, sign_inv_net="masked_gin", phi_out_dim=8,
                           sign_inv_layers=3, sign_inv_activation="relu", pe_aggregate="concat")


@pytest.mark.parametrize("pe_aggregate,edge_feat,readout", [("concat", True, "mean"), ("add", True, "sum"),
                                                             ("add", False, "mean")])
def test_gatedgcn_net_predictor(pe_aggregate, edge_feat, readout):
    """the GatedGCN predictor of GatedGCN_ZINC_LapPE_signinv_GIN_mask.json (gatedgcn_net.py:86-135,
    gatedgcn_layer.py:36-77) consuming the sign-invariant PE."""
    gg = ref_loader.gatedgcn_net()   # puts the stand-in dgl on sys.path
    import dgl

    torch.manual_seed(9)
    params = dict(GATEDGCN_NET_PARAMS, pe_aggregate=pe_aggregate, edge_feat=edge_feat, readout=readout)
    net = gg.GatedGCNNet(params)
    k = params["pos_enc_dim"]
    d = synth_batch(6, "zinc", seed=21, k_dgl=k)
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    sd = _leafify(_clone_sd(net))
    atoms, bonds = d.x[:, 0], d.edge_attr.reshape(-1)
    pe = net.sign_inv_net(g, d.pos_enc.unsqueeze(-1)).squeeze(-1)
    ref, _ = net(g, atoms, pe, bonds, None)
    pe_o = restate.masked_gin_deepsigns(d.pos_enc.unsqueeze(-1), d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph,
                                        {k2[len("sign_inv_net."):]: v for k2, v in sd.items() if k2.startswith("sign_inv_net.")},
                                        params["sign_inv_layers"], k).squeeze(-1)
    out = restate.gatedgcn_net(atoms, pe_o, bonds, d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph, sd,
                               params["L"], readout, edge_feat, pe_aggregate)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-5)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (out * w).sum().backward()
    want = {k2: v.grad for k2, v in net.named_parameters() if v.grad is not None}
    got = {k2: v.grad for k2, v in sd.items() if v.requires_grad and v.grad is not None and not k2.endswith(".eps")}
    assert set(want) == set(got)   # dgl GINConv eps (inside sign_inv_net) is a buffer
    assert_grads_close(got, want, 5e-5, "gatedgcn_net")
    for name, buf in net.named_buffers():   # BatchNorm running statistics of both streams
        if buf.is_floating_point():
            torch.testing.assert_close(sd[name], buf, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("method", ["sign_flip", "abs_val", "canonical", "none"])
def test_handle_lap_baselines(method):
    """The non-learned PE baselines of train_ZINC_graph_regression.py:12-47 (bit-exact: sign / abs only)."""
    import types

    tr = ref_loader.zinc_train_loop()   # puts the stand-in dgl on sys.path
    import dgl

    d = synth_batch(7, "zinc", seed=4, k_dgl=8)
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    model = types.SimpleNamespace(lap_method=method)
    torch.manual_seed(3)
    ref = tr.handle_lap(model, d.pos_enc.clone(), g, "cpu")
    torch.manual_seed(3)
    flip = torch.rand(d.pos_enc.size(1))
    flip[flip >= 0.5] = 1.0
    flip[flip < 0.5] = -1.0
    out = restate.handle_lap(d.pos_enc.clone(), d.num_nodes_per_graph, method, sign_flip=flip)
    assert torch.equal(out, ref)


PNA_NET_PARAMS = dict(num_atom_type=28, num_bond_type=4, hidden_dim=20, out_dim=20, in_feat_dropout=0.0, dropout=0.0, L=3,
                      readout="sum", graph_norm=True, batch_norm=True, residual=True, aggregators="mean max min std",
                      scalers="identity amplification attenuation", avg_d={"log": 1.1}, towers=5, divide_input_first=True,
                      divide_input_last=True, edge_feat=True, edge_dim=8, pretrans_layers=1, posttrans_layers=1, gru=False,
                      device="cpu", pe_init="lap_pe", lap_method="sign_inv", lap_lspe=False, use_lapeig_loss=False,
                      lambda_loss=1000, alpha_loss=1e-4, pos_enc_dim=6, sign_inv_net="masked_gin", phi_out_dim=8,
                      sign_inv_layers=3, sign_inv_activation="relu", pe_aggregate="concat")


@pytest.mark.parametrize("towers,divide,readout", [(5, True, "sum"), (1, True, "mean"), (2, True, "sum")])
def test_pna_net_predictor(towers, divide, readout):
    """SURVEY 8f rank 4 (oracle side only so far): the PNA predictor of PNA_ZINC_LapPE_signinv_GIN_mask.json
    (pna_net.py:116-167, pna_layer.py:16-153, pna_utils.py aggregators / scalers) consuming the sign-invariant PE.
    Reference quirk: divide_input=False cannot run (pna_layer.py:146 passes 5 arguments to PNATower.forward, which takes
    4), so only the divide_input=True form every shipped configuration uses is pinned."""
    pn = ref_loader.pna_net()   # puts the stand-in dgl on sys.path
    import dgl

    torch.manual_seed(11)
    params = dict(PNA_NET_PARAMS, towers=towers, divide_input_first=divide, divide_input_last=divide, readout=readout)
    net = pn.PNANet(params)
    k = params["pos_enc_dim"]
    d = synth_batch(6, "zinc", seed=22, k_dgl=k)
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    sd = _leafify(_clone_sd(net))
    atoms, bonds = d.x[:, 0], d.edge_attr.reshape(-1)
    n = torch.as_tensor(d.num_nodes_per_graph)
    snorm_n = (1.0 / n.float().sqrt()).repeat_interleave(n).unsqueeze(1)   # data/molecules.py collate: 1/sqrt(n_b) per node
    pe = net.sign_inv_net(g, d.pos_enc.unsqueeze(-1)).squeeze(-1)
    ref, _ = net(g, atoms, pe, bonds, snorm_n)
    pe_o = restate.masked_gin_deepsigns(d.pos_enc.unsqueeze(-1), d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph,
                                        {k2[len("sign_inv_net."):]: v for k2, v in sd.items() if k2.startswith("sign_inv_net.")},
                                        params["sign_inv_layers"], k).squeeze(-1)
    out = restate.pna_net(atoms, pe_o, bonds, d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph, snorm_n, sd,
                          params["L"], towers, params["avg_d"]["log"], readout, divide, divide)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-5)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (out * w).sum().backward()
    want = {k2: v.grad for k2, v in net.named_parameters() if v.grad is not None}
    got = {k2: v.grad for k2, v in sd.items() if v.requires_grad and v.grad is not None and not k2.endswith(".eps")}
    assert set(want) == set(got)
    assert_grads_close(got, want, 5e-5, "pna_net")


TRANSFORMER_NET_PARAMS = dict(num_atom_type=28, num_bond_type=4, hidden_dim=16, out_dim=16, n_heads=4, full_graph=False,
                              in_feat_dropout=0.0, dropout=0.0, L=3, readout="sum", batch_norm=True, layer_norm=True,
                              residual=True, edge_feat=True, device="cpu", pe_init="lap_pe", lap_method="sign_inv",
                              lap_lspe=False, use_lapeig_loss=False, lambda_loss=1, alpha_loss=1e-4, pos_enc_dim=6,
                              sign_inv_net="gin", phi_out_dim=4, sign_inv_layers=3, sign_inv_activation="relu",
                              pe_aggregate="concat")


@pytest.mark.parametrize("pe_aggregate,readout,n_heads", [("concat", "sum", 4), ("add", "mean", 2)])
def test_transformer_net_predictor(pe_aggregate, readout, n_heads):
    """SURVEY 8f rank 4 (oracle side only so far): the sparse graph Transformer of Transformer_ZINC_LapPE_signinv_GIN.json
    (transformer_net.py:89-150, transformer.py:117-301 with full_graph=False) consuming the sign-invariant PE.
    Reference quirk: TransformerNet never forwards `layer_norm` / `use_bias` to its layers, so there is no LayerNorm and
    the Q/K/E/V projections have no bias whatever net_params says (transformer_net.py:68-69)."""
    tn = ref_loader.transformer_net()   # puts the stand-in dgl on sys.path
    import dgl

    torch.manual_seed(13)
    params = dict(TRANSFORMER_NET_PARAMS, pe_aggregate=pe_aggregate, readout=readout, n_heads=n_heads)
    net = tn.TransformerNet(params)
    k = params["pos_enc_dim"]
    d = synth_batch(6, "zinc", seed=24, k_dgl=k)
    g = dgl.BatchedGraph(d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph)
    sd = _leafify(_clone_sd(net))
    atoms, bonds = d.x[:, 0], d.edge_attr.reshape(-1)
    pe = net.sign_inv_net(g, d.pos_enc.unsqueeze(-1)).squeeze(-1)
    ref, _ = net(g, atoms, pe, bonds, None)
    pe_o = restate.gin_deepsigns(d.pos_enc.unsqueeze(-1), d.edge_index[0], d.edge_index[1],
                                 {k2[len("sign_inv_net."):]: v for k2, v in sd.items() if k2.startswith("sign_inv_net.")},
                                 params["sign_inv_layers"], k).squeeze(-1)
    out = restate.transformer_net(atoms, pe_o, bonds, d.edge_index[0], d.edge_index[1], d.num_nodes_per_graph, sd,
                                  params["L"], n_heads, readout, pe_aggregate)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-5)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (out * w).sum().backward()
    want = {k2: v.grad for k2, v in net.named_parameters() if v.grad is not None}
    got = {k2: v.grad for k2, v in sd.items() if v.requires_grad and v.grad is not None and not k2.endswith(".eps")}
    assert set(want) == set(got)   # gamma (full-graph mixing weight) never reaches the output: no gradient on either side
    assert_grads_close(got, want, 5e-5, "transformer_net")

"""Batched device-side Laplacian EVD against torch.linalg.eigh of the same Laplacians on the CPU
(the reference's EVDTransform('sym'), Alchemy/sign_net/transform.py:7-23; Laplacian restated in synth.sym_laplacian)."""
import pytest
import torch

from signnet_basisnet_b200.synth import sym_laplacian, synth_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _per_graph(d):
    n = d.num_nodes_per_graph.tolist()
    off, out = 0, []
    for nb in n:
        m = (d.edge_index[0] >= off) & (d.edge_index[0] < off + nb)
        out.append((nb, d.edge_index[:, m] - off))
        off += nb
    return out


@pytest.mark.parametrize("shape,B,seed", [("zinc", 64, 3), ("alchemy", 50, 4), ("zinc", 1, 5)])
def test_laplacian_evd_matches_eigh(shape, B, seed):
    from signnet_basisnet_b200.ops import laplacian_evd

    d = synth_batch(B, shape, seed=seed)
    lam, vec = laplacian_evd(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    lam, vec = lam.cpu(), vec.cpu()
    o1 = o2 = 0
    for nb, ei in _per_graph(d):
        L = sym_laplacian(ei, nb, torch.float64)
        w, Q = torch.linalg.eigh(L)
        lg = lam[o1:o1 + nb].double()
        Vg = vec[o2:o2 + nb * nb].reshape(nb, nb).double()
        assert torch.all(lg[1:] >= lg[:-1]), "eigenvalues must be ascending (eigh order)"
        assert (lg - w).abs().max() <= 2e-5, f"eigenvalues off by {(lg - w).abs().max():.2e}"   # 1e-5 of |L| = 2, fp32 Jacobi
        assert (Vg.T @ Vg - torch.eye(nb, dtype=torch.float64)).abs().max() <= 2e-5, "V not orthonormal"
        assert (Vg @ torch.diag(lg) @ Vg.T - L).abs().max() <= 2e-5, "V diag(lambda) V^T != L"
        # eigenspace projectors (basis / sign free) of clusters separated from the rest by more than 1e-2:
        # fp32 eigenvector error ~ eps * |L| / gap <= 1e-4 there
        edges = [0] + [i for i in range(1, nb) if w[i] - w[i - 1] > 1e-2] + [nb]
        for a, b in zip(edges[:-1], edges[1:]):
            P_ref = Q[:, a:b] @ Q[:, a:b].T
            P_got = Vg[:, a:b] @ Vg[:, a:b].T
            assert (P_ref - P_got).abs().max() <= 1e-3, f"eigenspace [{a},{b}) differs by {(P_ref - P_got).abs().max():.2e}"
        o1 += nb
        o2 += nb * nb


def test_evd_feeds_signnet_like_the_cpu_transform():
    """The model is sign invariant, so on graphs with simple spectra the device EVD and the CPU eigh give the same
    SignNetGNN output (eval mode) - the drop-in property of the fused data path."""
    from signnet_basisnet_b200.ops import laplacian_evd
    from signnet_basisnet_b200.sign_net import SignNetGNN

    torch.manual_seed(0)
    d = synth_batch(24, "alchemy", seed=21)
    # keep graphs whose Laplacian spectrum is simple (gap > 1e-2): degenerate eigenspaces have no canonical basis
    keep, off = [], 0
    for g, (nb, ei) in enumerate(_per_graph(d)):
        w = torch.linalg.eigvalsh(sym_laplacian(ei, nb, torch.float64))
        if nb < 2 or float((w[1:] - w[:-1]).min()) > 1e-2:
            keep.append(g)
    if len(keep) < 2:
        pytest.skip("no simple-spectrum graphs in this sample")
    from signnet_basisnet_b200.ddp import shard_batch  # noqa: F401  (documented API; selection below is manual)
    sel = torch.tensor(keep)
    node_mask = torch.isin(d.batch, sel)
    remap = -torch.ones(d.num_graphs, dtype=torch.int64)
    remap[sel] = torch.arange(len(keep))
    new_id = torch.cumsum(node_mask.to(torch.int64), 0) - 1
    em = node_mask[d.edge_index[0]]
    import copy
    s = copy.copy(d)
    s.x, s.batch = d.x[node_mask], remap[d.batch[node_mask]]
    s.edge_index, s.edge_attr = new_id[d.edge_index[:, em]], d.edge_attr[em]
    s.num_graphs = len(keep)
    n = d.num_nodes_per_graph
    vptr = torch.cat([n.new_zeros(1), (n * n).cumsum(0)])
    s.eigen_values = d.eigen_values[node_mask]
    s.eigen_vectors = torch.cat([d.eigen_vectors[vptr[g]:vptr[g + 1]] for g in keep])
    s.num_nodes_per_graph = n[sel]
    model = SignNetGNN(6, 4, n_hid=16, n_out=3, nl_signnet=2, nl_gnn=2).to(DEV).eval()
    sg = s.to(DEV)
    with torch.no_grad():
        ref = model(sg)
        lam, vec = laplacian_evd(sg.edge_index, sg.batch, sg.num_graphs)
        sg.eigen_values, sg.eigen_vectors = lam, vec
        sg.__dict__.pop("_b200_graph_index", None)
        out = model(sg)
    assert (out - ref).abs().max() <= 2e-3 * ref.abs().max()


def test_lap_positional_encoding_dgl_convention():
    """DGL tree (molecules.py:148-181): trivial eigenvector dropped, next k kept, zero padded; columns match the CPU
    eigh up to sign wherever the eigenvalue is isolated."""
    from signnet_basisnet_b200.ops import lap_positional_encoding

    k = 8
    d = synth_batch(40, "alchemy", seed=9, k_dgl=k)
    pe = lap_positional_encoding(d.edge_index.to(DEV), d.batch.to(DEV), k, d.num_graphs).cpu()
    assert pe.shape == d.pos_enc.shape
    off = checked = 0
    for nb, ei in _per_graph(d):
        w = torch.linalg.eigvalsh(sym_laplacian(ei, nb, torch.float64))
        a, b = pe[off:off + nb], d.pos_enc[off:off + nb]
        m = min(k, nb - 1)
        assert a[:, m:].abs().max() == 0 if m < k else True
        for j in range(m):
            e = j + 1   # eigenvector index (the trivial one is dropped)
            gap = min(float(w[e] - w[e - 1]), float(w[e + 1] - w[e]) if e + 1 < nb else 1.0)
            if gap > 1e-2:
                err = min(float((a[:, j] - b[:, j]).abs().max()), float((a[:, j] + b[:, j]).abs().max()))
                assert err <= 1e-3, (j, err)
                checked += 1
        off += nb
    assert checked > 20

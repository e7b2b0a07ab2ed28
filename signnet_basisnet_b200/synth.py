"""Seeded synthetic graph batches of the shapes BASELINE.json names.

Everything here runs on the host with plain torch (it is the DataLoader's side of the boundary, cf. the reference's
EVD pre-transform Alchemy/sign_net/transform.py:7-23 and its DGL twin GraphPrediction/data/molecules.py:148-181);
identical tensors are fed to the oracle and to the CUDA path.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch


class Data(SimpleNamespace):
    """Duck-type of the torch_geometric `Batch` the reference's forward(data) consumes (sign_net.py:97,113;
    model.py:37-49): .x .edge_index .edge_attr .batch .eigen_values .eigen_vectors (+ .num_graphs)."""

    def to(self, device, non_blocking=False):
        out = Data()
        for k, v in self.__dict__.items():
            setattr(out, k, v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
        return out

    def pin_memory(self):
        out = Data()
        for k, v in self.__dict__.items():
            setattr(out, k, v.pin_memory() if torch.is_tensor(v) else v)
        return out


def _graph_sizes(B, shape, gen):
    if shape == "zinc":
        n = torch.round(23.2 + 4.6 * torch.randn(B, generator=gen)).clamp_(9, 37).long()
    elif shape == "alchemy":
        n = torch.randint(6, 13, (B,), generator=gen)
    else:
        raise ValueError(f"unknown shape {shape!r}")
    return n


def _molecule_like_edges(n, gen):
    """Random recursive tree + max(1, n//8) chords; symmetrised, de-duplicated, sorted by (src, dst)."""
    child = torch.arange(1, n)
    parent = (torch.rand(n - 1, generator=gen) * child).floor().long()
    n_chord = max(1, n // 8)
    a = torch.randint(0, n, (n_chord,), generator=gen)
    b = torch.randint(0, n, (n_chord,), generator=gen)
    keep = a != b
    src = torch.cat([child, parent, a[keep], b[keep]])
    dst = torch.cat([parent, child, b[keep], a[keep]])
    key = torch.unique(src * n + dst)
    return torch.stack([key // n, key % n])


def sym_laplacian(edge_index, n, dtype=torch.float32):
    """I - D^-1/2 A D^-1/2 with inf -> 0 (matches get_laplacian(..., 'sym'), transform.py:18-20)."""
    A = torch.zeros(n, n, dtype=dtype)
    A[edge_index[0], edge_index[1]] = 1.0
    deg = A.sum(1)
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    return torch.eye(n, dtype=dtype) - dis[:, None] * A * dis[None, :]


def synth_batch(B, shape="zinc", seed=0, k_dgl=None):
    """-> Data with PyG-convention ragged eigen data; if `k_dgl` is given also `.pos_enc [N, k_dgl]` in the DGL
    convention (eigvec 0 dropped, columns 1..k, zero-padded; molecules.py:159-177)."""
    gen = torch.Generator().manual_seed(seed)
    sizes = _graph_sizes(B, shape, gen)
    eis, evals, evecs, pes = [], [], [], []
    off = 0
    for n in sizes.tolist():
        ei = _molecule_like_edges(n, gen)
        L = sym_laplacian(ei, n)
        D, V = torch.linalg.eigh(L)
        eis.append(ei + off)
        evals.append(D)
        evecs.append(V.reshape(-1))
        if k_dgl is not None:
            pe = torch.zeros(n, k_dgl)
            m = min(k_dgl, n - 1)
            pe[:, :m] = V[:, 1 : 1 + m]
            pes.append(pe)
        off += n
    N = off
    edge_index = torch.cat(eis, 1)
    E = edge_index.shape[1]
    batch = torch.repeat_interleave(torch.arange(B), sizes)
    d = Data(edge_index=edge_index, batch=batch, eigen_values=torch.cat(evals), eigen_vectors=torch.cat(evecs),
             num_graphs=B, num_nodes_per_graph=sizes)
    if shape == "zinc":
        d.x = torch.randint(0, 28, (N, 1), generator=gen)
        d.edge_attr = torch.randint(1, 4, (E,), generator=gen)
    else:
        d.x = torch.rand(N, 6, generator=gen)
        d.edge_attr = torch.rand(E, 4, generator=gen)
    if k_dgl is not None:
        d.pos_enc = torch.cat(pes)
    return d


def grid_graph(h, w):
    """h x w 4-neighbour grid, both directions, sorted by (src, dst) (cfg 1 / cfg 5 single-graph workloads)."""
    idx = torch.arange(h * w).view(h, w)
    e = torch.cat([torch.stack([idx[:, :-1].reshape(-1), idx[:, 1:].reshape(-1)]),
                   torch.stack([idx[:-1, :].reshape(-1), idx[1:, :].reshape(-1)])], 1)
    e = torch.cat([e, e.flip(0)], 1)
    key = torch.unique(e[0] * (h * w) + e[1])
    return torch.stack([key // (h * w), key % (h * w)])

"""Stand-alone bookkeeping ops of the path with the reference's call signatures (rows a1/a2 of SURVEY.md §8)."""
from __future__ import annotations

import torch

from ._lib import counted_call as _call, ptr as _p
from .layout import GraphIndex


def to_dense_list_EVD(eigS, eigV, batch, return_mask=False, num_graphs=None):
    """Alchemy/sign_net/transform.py:52-61 — ragged per-graph EVD -> (eigS_dense [N,Nmax], eigV_dense [N,Nmax]),
    bit-exact, without materialising [B,Nmax,Nmax].  `return_mask` adds mask_full [N,Nmax] (sign_net.py:100-102)."""
    for t, name in ((eigS, "eigS"), (eigV, "eigV")):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise ValueError(f"{name} must be a CUDA float32 tensor")
    dev = batch.device
    N = batch.numel()
    empty_edges = torch.empty(2, 0, dtype=torch.int64, device=dev)
    gi = GraphIndex(empty_edges, batch, num_graphs)
    sl = gi.slots_all(4)
    if eigV.numel() != sl.vec_total:
        raise ValueError(f"eigV has {eigV.numel()} entries, batch needs sum n_b^2 = {sl.vec_total}")
    nmax = sl.nmax
    S = torch.empty(N, nmax, dtype=torch.float32, device=dev)
    V = torch.empty(N, nmax, dtype=torch.float32, device=dev)
    M = torch.empty(N, nmax, dtype=torch.bool, device=dev) if return_mask else None
    _call("sb_dense_list_evd", _p(eigS.contiguous()), _p(eigV.contiguous()), _p(gi.batch), _p(gi.graph_ptr),
          _p(sl.vec_ptr), N, nmax, _p(S), _p(V), _p(M))
    return (S, V, M) if return_mask else (S, V)


def laplacian_evd(edge_index, batch, num_graphs=None, graph_index=None):
    """Batched `EVDTransform('sym')` on the device (Alchemy/sign_net/transform.py:7-23): per graph of the batch,
    eigh of L = I - D^-1/2 A D^-1/2 (A = to_undirected(edge_index), self loops removed).  Returns
    (eigen_values [N] ascending per graph, eigen_vectors [sum n_b^2] row-major V[node, eig]) - exactly the fields the
    PyG trees' `forward(data)` reads - plus nothing on the host: one warp per graph, parallel cyclic Jacobi in shared
    memory (csrc/evd.cu).  Eigenvector signs (and bases of degenerate eigenspaces) are arbitrary, as with LAPACK."""
    gi = graph_index or GraphIndex(edge_index, batch, num_graphs)
    sl = gi.slots_all(4)
    dev = gi.device
    evals = torch.empty(gi.N, dtype=torch.float32, device=dev)
    evecs = torch.empty(max(sl.vec_total, 1), dtype=torch.float32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    _call("sb_laplacian_evd", _p(gi.graph_ptr), _p(gi.in_ptr), _p(gi.in_src), _p(sl.vec_ptr), gi.B, sl.nmax, _p(evals),
          _p(evecs), _p(flags))
    return evals, evecs[:sl.vec_total]


def lap_positional_encoding(edge_index, batch, pos_enc_dim, num_graphs=None):
    """DGL-tree convention (GraphPrediction/data/molecules.py:148-181): eigenvectors of the sym-normalised Laplacian
    sorted by eigenvalue, the trivial one dropped, the next `pos_enc_dim` kept, zero padded -> pos_enc [N, pos_enc_dim]."""
    gi = GraphIndex(edge_index, batch, num_graphs)
    _, evecs = laplacian_evd(edge_index, batch, num_graphs, graph_index=gi)
    _, V = to_dense_list_EVD(torch.zeros(gi.N, dtype=torch.float32, device=gi.device), evecs, gi.batch,
                             num_graphs=gi.B)
    pe = torch.zeros(gi.N, pos_enc_dim, dtype=torch.float32, device=gi.device)
    take = min(pos_enc_dim, max(V.shape[1] - 1, 0))
    pe[:, :take] = V[:, 1:1 + take]
    return pe

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

"""Per-batch device bookkeeping for the SignNet hot path: graph offsets, stable CSR/CSC, ragged slot-row layout.

Replaces the index math the reference redoes inside every forward with torch_scatter / boolean masks
(Alchemy/sign_net/transform.py:26-61, sign_net.py:100-102; GraphPrediction/layers/deepsigns.py:66-78).  All arrays are
built by the integer kernels of csrc/bookkeeping.cu; one small device->host copy (the `summary` vector) gives the
host the sizes it needs to allocate (the reference has two such syncs, transform.py:28,31).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import counted_call as _call, ptr as _p

FLAG_MESSAGES = {
    1: "`batch` must be sorted (non-decreasing)",
    2: "`batch` holds a value outside [0, num_graphs)",
    4: "`edge_index` holds a node id outside [0, N)",
    8: "`edge_index` joins nodes of different graphs",
}


def _check_index(t, name, ndim):
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.int64 and t.dim() == ndim):
        raise ValueError(f"{name} must be a CUDA int64 tensor of rank {ndim}")
    return t.contiguous()


def pad4(d: int) -> int:
    return (d + 3) // 4 * 4


class GraphIndex:
    """graph_ptr + CSR (by destination) + CSC (by source) of a batched graph.  Independent of k."""

    def __init__(self, edge_index, batch, num_graphs=None, allow_cross_graph=False):
        self.edge_index = _check_index(edge_index, "edge_index", 2)
        self.batch = _check_index(batch, "batch", 1)
        if self.edge_index.shape[0] != 2:
            raise ValueError("edge_index must have shape [2, E]")
        dev = self.batch.device
        self.device = dev
        self.N = int(self.batch.numel())
        self.E = int(self.edge_index.shape[1])
        if num_graphs is None:
            num_graphs = int(self.batch[-1].item()) + 1 if self.N > 0 else 0  # same host sync as transform.py:28
        self.B = int(num_graphs)
        N, E, B = self.N, self.E, self.B
        i32 = dict(dtype=torch.int32, device=dev)
        self.flags = torch.zeros(1, **i32)
        self.graph_ptr = torch.empty(B + 1, **i32)
        _call("sb_graph_ptr", _p(self.batch), N, B, _p(self.graph_ptr), _p(self.flags))
        self.in_ptr = torch.empty(N + 1, **i32)
        self.in_src = torch.empty(max(E, 1), **i32)
        self.in_eid = torch.empty(max(E, 1), **i32)
        self.out_ptr = torch.empty(N + 1, **i32)
        self.out_dst = torch.empty(max(E, 1), **i32)
        self.out_eid = torch.empty(max(E, 1), **i32)
        ws_ints = 2 * (N + 1) + 2 * ((N + 1 + 4095) // 4096) + 8
        ws = torch.empty(ws_ints, **i32)
        _call("sb_build_csr", _p(self.edge_index), E, N, _p(self.batch), _p(self.in_ptr), _p(self.in_src),
              _p(self.in_eid), _p(self.out_ptr), _p(self.out_dst), _p(self.out_eid), _p(ws), ws_ints, _p(self.flags))
        # packed neighbour words of both CSRs for the TMA aggregate (4 local ids per node in one int32)
        self.in_pack = torch.empty(max(N, 1), **i32)
        self.out_pack = torch.empty(max(N, 1), **i32)
        _call("sb_pack_neighbours", _p(self.batch), _p(self.graph_ptr), _p(self.in_ptr), _p(self.in_src), N,
              _p(self.in_pack))
        _call("sb_pack_neighbours", _p(self.batch), _p(self.graph_ptr), _p(self.out_ptr), _p(self.out_dst), N,
              _p(self.out_pack))
        self._allow_cross = allow_cross_graph
        self._checked = False
        self._slots = {}

    def tensors(self):
        out = [v for v in self.__dict__.values() if torch.is_tensor(v)]
        for sl in self._slots.values():
            out += [v for v in sl.__dict__.values() if torch.is_tensor(v)]
        return out

    def wait_ready(self):
        """Make torch's current stream wait for the stream this index was built on (prepare_batch); no-op otherwise."""
        ev = self.__dict__.get("ready")
        if ev is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for t in self.tensors():     # allocated on the builder's stream, consumed on this one
                t.record_stream(cur)
            self.ready = None

    def check(self, flags_host=None):
        """Raise ValueError for malformed inputs (mirrors the reference's implicit index errors)."""
        if self._checked:
            return
        f = int(self.flags.item()) if flags_host is None else int(flags_host)
        if self._allow_cross:
            f &= ~8
        if f:
            raise ValueError("; ".join(m for b, m in FLAG_MESSAGES.items() if f & b))
        self._checked = True

    def slots(self, k: int, masked: bool = True, ld: int = 128) -> "SlotLayout":
        """Slot-row layout for k eigenvector slots; cached per (k, masked, tile_rows)."""
        tile_rows = _lib.lib().sb_gin_agg_tile_rows(int(ld))
        key = (int(k), bool(masked), tile_rows)
        if key not in self._slots:
            base = next((v for kk, v in self._slots.items() if kk[:2] == key[:2]), None)
            self._slots[key] = SlotLayout(self, int(k), bool(masked), tile_rows, base)
        return self._slots[key]


    def slots_all(self, ld: int = 128) -> "SlotLayout":
        """Layout with k = N_max of the batch (the PyG trees' setting, transform.py:31): every graph keeps all of its
        n_b eigenvectors.  One layout kernel + one host sync."""
        tile_rows = _lib.lib().sb_gin_agg_tile_rows(int(ld))
        hit = next((v for kk, v in self._slots.items() if kk[1] and kk[2] == tile_rows and v.k == v.nmax), None)
        if hit is not None:
            return hit
        sl = SlotLayout(self, 1 << 20, True, tile_rows, None)
        sl.k = max(sl.nmax, 1)
        self._slots[(sl.k, True, tile_rows)] = sl
        return sl


class SlotLayout:
    """row(b, j, i) = row_ptr[b] + j*n_b + i for slot j < k_b (= min(n_b, k) if masked else k)."""

    def __init__(self, gi: GraphIndex, k: int, masked: bool, tile_rows: int, base: "SlotLayout | None" = None):
        self.gi, self.k, self.masked, self.tile_rows = gi, k, masked, tile_rows
        dev, B = gi.device, gi.B
        self.unit_ptr = torch.empty(B + 1, dtype=torch.int32, device=dev)
        if base is not None:  # same rows, different aggregate tile size
            self.row_ptr, self.vec_ptr = base.row_ptr, base.vec_ptr
            self.R, self.nmax, self.kmax, self.vec_total = base.R, base.nmax, base.kmax, base.vec_total
            _call("sb_agg_units", _p(gi.graph_ptr), B, k, int(masked), max(tile_rows, 1), _p(self.unit_ptr))
            self.oversize = 0 if (tile_rows >= self.nmax and tile_rows > 0) else 1
            self._unit_desc()
            return
        self.row_ptr = torch.empty(B + 1, dtype=torch.int64, device=dev)
        self.vec_ptr = torch.empty(B + 1, dtype=torch.int64, device=dev)
        summary = torch.zeros(8, dtype=torch.int64, device=dev)
        _call("sb_slot_layout", _p(gi.graph_ptr), B, k, int(masked), max(tile_rows, 1), _p(self.row_ptr),
              _p(self.vec_ptr), _p(self.unit_ptr), _p(summary))
        host = torch.cat([summary, gi.flags.to(torch.int64)]).cpu()  # the one host sync of the layout
        gi.check(host[8])
        self.R, self.nmax, self.kmax, self.vec_total = (int(host[i]) for i in range(4))
        self.oversize = int(host[5]) if tile_rows > 0 else 1
        self._unit_desc()

    def _unit_desc(self):
        """One 48-byte record per aggregate tile (sb_agg_unit_desc); #tiles <= N (masked: k_b <= n_b) or B*k."""
        gi = self.gi
        cap = max(1, min(self.R, gi.N if self.masked else gi.B * self.k))
        self.unit_desc = torch.empty(cap, 12, dtype=torch.int32, device=gi.device)
        _call("sb_agg_unit_desc", _p(gi.graph_ptr), _p(self.row_ptr), _p(self.unit_ptr), _p(gi.in_ptr), _p(gi.out_ptr),
              gi.B, self.k, int(self.masked), max(self.tile_rows, 1), _p(self.unit_desc), cap)

    @property
    def use_generic_agg(self) -> bool:
        return self.oversize > 0 or self.tile_rows <= 0


def prepare_batch(data, ld: int = 128, stream=None, k=None, masked=True):
    """Build the per-batch bookkeeping (GraphIndex + the slot-row layout for row stride `ld`) ahead of the step and
    attach it to `data` (`data._b200_graph_index`, which the modules' forward(data) picks up).

    With `stream` (a side stream on which `data`'s tensors are complete - e.g. the stream of the prefetching H2D copy)
    the integer kernels AND the layout's one device->host read run there, one step ahead of the compute stream: the
    host never waits for the previous step's backward in the middle of a step (at 128 graphs per GPU that wait made the
    forward host-bound, scripts/host_probe.py).  The consumer calls GraphIndex.wait_ready() (done by the modules).
    k=None selects the PyG trees' layout (k = N_max of the batch); otherwise the (k, masked) layout of the DGL trees."""
    def build():
        gi = GraphIndex(data.edge_index, data.batch, getattr(data, "num_graphs", None))
        if k is None:
            gi.slots_all(ld)
        else:
            gi.slots(k, masked, ld)
        return gi

    if stream is None:
        gi = build()
    else:
        with torch.cuda.stream(stream):
            gi = build()
            gi.ready = torch.cuda.Event()
            gi.ready.record(stream)
    data._b200_graph_index = gi
    return gi

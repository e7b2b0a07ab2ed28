"""Graph-sharded data parallelism for the SignNet path: one process per GPU, every rank runs the full model on its own
shard of graphs (no halo, no exchange in the forward: graphs are independent), BatchNorm statistics stay per replica
(the reference has no SyncBN), and ONE all-reduce of a flat fp32 gradient buffer per step averages the gradients over
NVLink/NVSwitch.  Parameters the step never touched (e.g. the final norm MaskedMLP allocates but
skips, masked_layers.py:43,61) contribute zeros, so every rank reduces the same layout.

The reference is single-process (no torch.distributed call site anywhere); this file is the multi-GPU row of the hot
path, not a port of anything.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_graphs(num_graphs: int, world: int, rank: int):
    """Contiguous, size-balanced split of graph ids [0, num_graphs) -> (start, stop) for `rank`."""
    base, rem = divmod(num_graphs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(data, world: int, rank: int):
    """Slice a PyG-style batch (sorted `batch`, edges grouped by graph) to this rank's graphs and re-base node ids."""
    from .synth import Data

    n = torch.bincount(data.batch, minlength=int(data.num_graphs))
    g0, g1 = shard_graphs(int(data.num_graphs), world, rank)
    node_ptr = torch.cat([n.new_zeros(1), n.cumsum(0)])
    vec_ptr = torch.cat([n.new_zeros(1), (n * n).cumsum(0)])
    a, b = int(node_ptr[g0]), int(node_ptr[g1])
    emask = (data.edge_index[0] >= a) & (data.edge_index[0] < b)
    out = Data(x=data.x[a:b], edge_index=data.edge_index[:, emask] - a, batch=data.batch[a:b] - g0,
               eigen_values=data.eigen_values[a:b], eigen_vectors=data.eigen_vectors[int(vec_ptr[g0]):int(vec_ptr[g1])],
               num_graphs=g1 - g0)
    if getattr(data, "edge_attr", None) is not None:
        out.edge_attr = data.edge_attr[emask]
    if getattr(data, "y", None) is not None:
        out.y = data.y[g0:g1]
    return out


class FlatGradAllReduce:
    """Average gradients across ranks with a single all-reduce of one flat fp32 buffer."""

    def __init__(self, module: torch.nn.Module, world: int | None = None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def broadcast_parameters(self, src: int = 0):
        """Make every replica start from rank `src`'s weights and buffers."""
        if self.world == 1:
            return
        with torch.no_grad():
            flat = torch.cat([p.detach().reshape(-1) for p in self.params])
            dist.broadcast(flat, src)
            off = 0
            for p in self.params:
                p.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()

    def allreduce(self):
        """Pack (zero-filling untouched parameters) -> all_reduce(sum) -> scale by 1/world -> hand back as .grad."""
        with torch.no_grad():
            self.flat.zero_()
            have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
            if self.world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
                self.flat.mul_(1.0 / self.world)
            for v, p in zip(self.views, self.params):
                p.grad = v
        return self.flat

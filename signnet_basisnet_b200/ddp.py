"""Graph-sharded data parallelism for the SignNet path: one process per GPU, every rank runs the full model on its own
shard of graphs (no halo, no exchange in the forward: graphs are independent), BatchNorm statistics stay per replica
(the reference has no SyncBN), and the gradients are averaged over NVLink/NVSwitch by all-reducing ONE flat fp32
buffer, cut into a few contiguous buckets that are launched on a side stream as soon as their gradients exist, so the
collective overlaps the rest of the backward.

What is reduced.  Parameters the step never touches - the final norm MaskedMLP allocates but skips
(masked_layers.py:43,61), and in the GINESignNetPyG tree phi.edge_encoders, eigen_encoder1/2, rho.pos_encoder and nine
of the ten DiscreteEncoder tables (core/sign_net.py:22,54,90; quirk v) - never receive a gradient; they are found on
the first step (`p.grad is None` after backward, checked to be the same set on every rank), keep `grad = None` exactly
as in single-process training (an optimizer skips them) and are not part of the payload: 4.4 MB instead of 38.4 MB for
the cfg 4 model.

The reference is single-process (no torch.distributed call site anywhere); this file is the multi-GPU row of the hot
path, not a port of anything.
"""
from __future__ import annotations

import os

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
    """Average gradients across ranks through one flat fp32 buffer, reduced in `buckets` overlapped pieces.

    Usage per step:  zero grads (any way: None, zeros, or not at all when accumulating) -> forward -> backward ->
    `allreduce()`.  After `allreduce()` every live parameter's `.grad` is a view of the flat buffer holding the
    average over ranks; untouched parameters keep `.grad = None`.

    Step 1 runs un-overlapped: it discovers which parameters receive gradients and in which order (hooks), checks that
    every rank found the same set, lays the flat buffer out in that order and cuts it into buckets.  From step 2 on a
    bucket is packed and all-reduced on a side stream the moment its last gradient has been accumulated.
    """

    def __init__(self, module: torch.nn.Module, world: int | None = None, buckets: int = 4, overlap: bool | None = None):
        self.module = module
        if overlap is None:
            overlap = os.environ.get("SB_DDP_OVERLAP", "1") != "0"
        self.overlap = bool(overlap)
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.n_buckets = max(1, int(buckets))
        self.device = self.params[0].device
        self.flat = None            # built after the first backward
        self.views = {}             # param index -> view of self.flat
        self.live, self.dead = [], []
        self._order = []            # param indices in the order their gradients were accumulated (first step)
        self._bucket_of, self._buckets = {}, []
        self._pending, self._works = [], []
        self._side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._avg_native = self.device.type == "cuda"   # NCCL has ReduceOp.AVG, gloo does not
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]
        self.last_payload_bytes = 0

    # ------------------------------------------------------------------------------------------------- parameters
    def broadcast_parameters(self, src: int = 0):
        """Make every replica start from rank `src`'s parameters AND floating-point buffers (BatchNorm running
        statistics); integer buffers (num_batches_tracked) are broadcast too."""
        if self.world == 1:
            return
        with torch.no_grad():
            tensors = list(self.module.parameters()) + list(self.module.buffers())
            fl = [t for t in tensors if t.is_floating_point()]
            if fl:
                flat = torch.cat([t.detach().reshape(-1).float() for t in fl])
                dist.broadcast(flat, src)
                off = 0
                for t in fl:
                    t.copy_(flat[off:off + t.numel()].view_as(t))
                    off += t.numel()
            for t in tensors:
                if not t.is_floating_point():
                    dist.broadcast(t, src)

    # ------------------------------------------------------------------------------------------------------ hooks
    def _make_hook(self, i):
        def hook(p):
            if self.flat is None:
                self._order.append(i)
                return
            b = self._bucket_of.get(i)
            if b is None:
                return                  # reported by allreduce(): a parameter thought dead received a gradient
            self._pending[b] -= 1
            if self._pending[b] == 0 and self.world > 1 and self.overlap:
                self._launch(b)
        return hook

    def _is_view(self, i):
        g = self.params[i].grad
        return g is not None and g.data_ptr() == self.views[i].data_ptr()

    def _launch(self, b):
        """Pack bucket b (gradients that are not already the flat views) and start its all-reduce."""
        lo, hi, idx = self._buckets[b]
        with torch.no_grad():
            if self._side is not None and self.overlap:
                ev = torch.cuda.Event()
                ev.record()                     # the backward stream: everything this bucket needs has been enqueued
                ctx = torch.cuda.stream(self._side)
                self._side.wait_event(ev)
            else:
                ctx = _Null()
            with ctx:
                dst, src = [], []
                for i in idx:
                    g = self.params[i].grad
                    if g is None:
                        self.views[i].zero_()   # a live parameter skipped this step: contributes zeros
                    elif not self._is_view(i):
                        dst.append(self.views[i])
                        src.append(g)
                if dst:
                    torch._foreach_copy_(dst, src)
                piece = self.flat[lo:hi]
                op = dist.ReduceOp.AVG if self._avg_native else dist.ReduceOp.SUM
                self._works.append(dist.all_reduce(piece, op=op, async_op=True))

    # ------------------------------------------------------------------------------------------------- first step
    def _build(self):
        got = [p.grad is not None for p in self.params]
        if self.world > 1:   # every rank must reduce the same layout
            m = torch.tensor([1.0 if g else 0.0 for g in got], device=self.device)
            lo, hi = m.clone(), m.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if not torch.equal(lo, hi):
                raise RuntimeError("FlatGradAllReduce: ranks disagree on which parameters receive gradients")
        seen = set()
        order = [i for i in self._order if got[i] and not (i in seen or seen.add(i))]
        order += [i for i, g in enumerate(got) if g and i not in seen]      # (gradients set without the hook firing)
        self.live, self.dead = order, [i for i, g in enumerate(got) if not g]
        total = sum(self.params[i].numel() for i in order)
        self.flat = torch.zeros(max(total, 1), dtype=torch.float32, device=self.device)
        off, cuts, target, nb = 0, [], total / self.n_buckets, 1
        self._buckets, cur, lo = [], [], 0
        for i in order:
            n = self.params[i].numel()
            self.views[i] = self.flat[off:off + n].view_as(self.params[i])
            cur.append(i)
            off += n
            if off >= target * nb and nb < self.n_buckets:
                self._buckets.append((lo, off, cur))
                cur, lo, nb = [], off, nb + 1
        if cur:
            self._buckets.append((lo, off, cur))
        self._bucket_of = {i: b for b, (_, _, idx) in enumerate(self._buckets) for i in idx}
        self.last_payload_bytes = 4 * total

    def _arm(self):
        self._pending = [len(idx) for _, _, idx in self._buckets]
        self._works = []

    # ----------------------------------------------------------------------------------------------------- public
    def allreduce(self):
        """Finish the step's gradient exchange; returns the flat buffer (averaged gradients of the live parameters)."""
        with torch.no_grad():
            first = self.flat is None
            if first:
                self._build()
                self._arm()
            else:
                for i in self.dead:
                    if self.params[i].grad is not None:
                        raise RuntimeError("FlatGradAllReduce: a parameter that had no gradient on the first step "
                                           "received one now; rebuild the reducer (the flat layout is fixed)")
            if self.world > 1:
                for b in range(len(self._buckets)):     # buckets whose hooks did not complete (first step; skipped params)
                    if first or self._pending[b] > 0:
                        self._launch(b)
                for w in self._works:
                    w.wait()                            # current stream waits for the collective
                if self._side is not None:
                    torch.cuda.current_stream().wait_stream(self._side)
                if not self._avg_native:
                    self.flat.mul_(1.0 / self.world)
            else:
                dst, src = [], []
                for i in self.live:
                    g = self.params[i].grad
                    if g is None:
                        self.views[i].zero_()
                    elif not self._is_view(i):
                        dst.append(self.views[i])
                        src.append(g)
                if dst:
                    torch._foreach_copy_(dst, src)
            for i in self.live:                         # only now may the step's own gradient tensors be released
                self.params[i].grad = self.views[i]
            self._arm()
        return self.flat

    def bucket_bytes(self):
        return [4 * (hi - lo) for lo, hi, _ in self._buckets]


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

"""TEST INFRASTRUCTURE ONLY — minimal stand-in for DGL (unpinned by the reference; APIs imply 0.5–0.6.x) so that
GraphPrediction/layers/{deepsigns,gnns,mlp}.py import unmodified.  Published semantics restated:
dgl.nn.pytorch.GINConv(apply_func, 'sum', init_eps=0, learn_eps=False):
    rst = (1 + eps) * feat + sum_{u->v} feat_u ;  return apply_func(rst) ; eps is a buffer unless learn_eps."""
import torch
from . import nn  # noqa: F401


class BatchedGraph:
    """Object exposing what deepsigns.py / gnns.py touch on a DGLGraph: edges() and batch_num_nodes()."""

    def __init__(self, src, dst, num_nodes_per_graph):
        self.src, self.dst = src, dst
        self._bnn = torch.as_tensor(num_nodes_per_graph)
        self.ndata, self.edata = {}, {}

    def edges(self):
        return self.src, self.dst

    def batch_num_nodes(self):
        return self._bnn

    def num_nodes(self):
        return int(self._bnn.sum())


def _segments(g):
    n = g.batch_num_nodes()
    return torch.repeat_interleave(torch.arange(n.numel()), n), n


def sum_nodes(g, key):
    """dgl.sum_nodes: per-graph sum of a node feature (gin_net.py:127-134 readout)."""
    seg, n = _segments(g)
    x = g.ndata[key]
    return torch.zeros(n.numel(), *x.shape[1:], dtype=x.dtype).index_add_(0, seg, x)


def mean_nodes(g, key):
    seg, n = _segments(g)
    s = sum_nodes(g, key)
    return s / n.to(s.dtype).clamp(min=1).reshape(-1, *([1] * (s.dim() - 1)))


def max_nodes(g, key):
    seg, n = _segments(g)
    x = g.ndata[key]
    return torch.stack([x[seg == b].max(0).values for b in range(n.numel())])

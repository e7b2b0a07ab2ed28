import torch
from . import glob  # noqa: F401
from .glob import SetTransformerEncoder  # noqa: F401


class GINConv(torch.nn.Module):
    def __init__(self, apply_func, aggregator_type, init_eps=0, learn_eps=False):
        super().__init__()
        assert aggregator_type == "sum"
        self.apply_func = apply_func
        if learn_eps:
            self.eps = torch.nn.Parameter(torch.FloatTensor([init_eps]))
        else:
            self.register_buffer("eps", torch.FloatTensor([init_eps]))

    def forward(self, graph, feat):
        src, dst = graph.edges()
        neigh = torch.zeros_like(feat).index_add_(0, dst, feat.index_select(0, src))
        rst = (1 + self.eps) * feat + neigh
        if self.apply_func is not None:
            rst = self.apply_func(rst)
        return rst


class _Unused(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


GraphConv = GATConv = _Unused

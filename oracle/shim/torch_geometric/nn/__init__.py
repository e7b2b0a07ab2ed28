"""Stand-ins for torch_geometric.nn 2.0.1 ops used on the path.

Published semantics restated (PyG 2.0.1, conv/gin_conv.py):
  GINConv.forward:  out = propagate(edge_index, x) = scatter_add(x.index_select(-2, edge_index[0]), edge_index[1],
                    dim=-2, dim_size=N);  out += (1 + eps) * x;  return nn(out)
  GINEConv.message: (x_j + edge_attr).relu()
  eps: Parameter([initial]) when train_eps else buffer; reset_parameters() resets nn and refills eps.
MessagePassing default node_dim = -2, so inputs of shape [k, N, d] aggregate over N.
Call sites: masked_layers.py:70,75 ; pyg_gnn_wrapper.py:11,23.
"""
import torch
from .inits import reset


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", node_dim=-2, **kwargs):
        super().__init__()
        self.aggr, self.node_dim = aggr, node_dim


def _neighbour_sum(msg, dst, like):
    out = torch.zeros_like(like)
    return out.index_add_(-2, dst, msg)


class GINConv(MessagePassing):
    def __init__(self, nn, eps=0.0, train_eps=False, **kwargs):
        super().__init__(aggr="add")
        self.nn, self.initial_eps = nn, eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))
        self.reset_parameters()

    def reset_parameters(self):
        reset(self.nn)
        self.eps.data.fill_(self.initial_eps)

    def forward(self, x, edge_index, size=None):
        out = _neighbour_sum(x.index_select(-2, edge_index[0]), edge_index[1], x)
        out += (1 + self.eps) * x
        return self.nn(out)


class GINEConv(MessagePassing):
    def __init__(self, nn, eps=0.0, train_eps=False, edge_dim=None, **kwargs):
        super().__init__(aggr="add")
        self.nn, self.initial_eps = nn, eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))
        self.lin = None
        self.reset_parameters()

    def reset_parameters(self):
        reset(self.nn)
        self.eps.data.fill_(self.initial_eps)

    def forward(self, x, edge_index, edge_attr=None, size=None):
        msg = (x.index_select(-2, edge_index[0]) + edge_attr).relu()
        out = _neighbour_sum(msg, edge_index[1], x)
        out += (1 + self.eps) * x
        return self.nn(out)


class _Unused(torch.nn.Module):
    """Constructible placeholder for convs the hot path never selects (GAT/GCN)."""

    def __init__(self, *a, **k):
        super().__init__()

    def reset_parameters(self):
        pass

    def forward(self, *a, **k):
        raise NotImplementedError("off the SignNet hot path")


GATConv = GCNConv = ARMAConv = ChebConv = APPNP = _Unused   # baselines of LearningFilters/models.py: import-only


def global_add_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else size
    return x.new_zeros((size,) + x.shape[1:]).index_add_(0, batch, x)

"""Stand-ins for torch_geometric.utils used by Alchemy/sign_net/transform.py:18-19 (EVD pre-transform)."""
import torch


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros(n, dtype=dtype if dtype is not None else torch.float)
    return out.scatter_add_(0, index, torch.ones_like(index, dtype=out.dtype))


def to_undirected(edge_index, num_nodes=None):
    row, col = edge_index
    row, col = torch.cat([row, col]), torch.cat([col, row])
    n = int(max(row.max(), col.max())) + 1 if num_nodes is None else num_nodes
    key = torch.unique(row * n + col)  # coalesce: sorted, duplicate-free
    return torch.stack([key // n, key % n])


def get_laplacian(edge_index, edge_weight=None, normalization=None, dtype=None, num_nodes=None):
    row, col = edge_index
    keep = row != col  # remove self loops
    row, col = row[keep], col[keep]
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    w = torch.ones(row.numel(), dtype=dtype if dtype is not None else torch.float) if edge_weight is None else edge_weight[keep]
    deg = torch.zeros(n, dtype=w.dtype).scatter_add_(0, row, w)
    loop = torch.arange(n)
    if normalization is None:
        ei = torch.cat([torch.stack([row, col]), torch.stack([loop, loop])], 1)
        return ei, torch.cat([-w, deg])
    if normalization == "sym":
        dis = deg.pow(-0.5)
        dis.masked_fill_(dis == float("inf"), 0)
        w = dis[row] * w * dis[col]
        ei = torch.cat([torch.stack([row, col]), torch.stack([loop, loop])], 1)
        return ei, torch.cat([-w, torch.ones(n, dtype=w.dtype)])
    if normalization == "rw":
        dinv = 1.0 / deg
        dinv.masked_fill_(dinv == float("inf"), 0)
        ei = torch.cat([torch.stack([row, col]), torch.stack([loop, loop])], 1)
        return ei, torch.cat([-(dinv[row] * w), torch.ones(n, dtype=w.dtype)])
    raise ValueError(normalization)


def add_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    """Import-only for LearningFilters/models.py; semantics of PyG 2.0.1 restated for completeness."""
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    loop = torch.arange(n, dtype=edge_index.dtype)
    ei = torch.cat([edge_index, torch.stack([loop, loop])], dim=1)
    if edge_weight is not None:
        edge_weight = torch.cat([edge_weight, edge_weight.new_full((n,), fill_value)])
    return ei, edge_weight

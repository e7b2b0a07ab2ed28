"""TEST INFRASTRUCTURE ONLY — minimal stand-in for DGL (unpinned by the reference; APIs imply 0.5–0.6.x) so that
GraphPrediction/layers/{deepsigns,gnns,mlp,gatedgcn_layer}.py and nets/ZINC_graph_regression/{gin,gatedgcn}_net.py import
unmodified.  Published semantics restated:
dgl.nn.pytorch.GINConv(apply_func, 'sum', init_eps=0, learn_eps=False):
    rst = (1 + eps) * feat + sum_{u->v} feat_u ;  return apply_func(rst) ; eps is a buffer unless learn_eps."""
import torch
from . import function  # noqa: F401
from . import nn  # noqa: F401


class BatchedGraph:
    """Object exposing what deepsigns.py / gnns.py touch on a DGLGraph: edges() and batch_num_nodes()."""

    def __init__(self, src, dst, num_nodes_per_graph):
        self.src, self.dst = src, dst
        self._bnn = torch.as_tensor(num_nodes_per_graph)
        self.ndata, self.edata = {}, {}

    def edges(self, form="uv"):
        if form == "eid":
            return torch.arange(self.src.numel())
        return self.src, self.dst

    def batch_num_nodes(self):
        return self._bnn

    def num_nodes(self):
        return int(self._bnn.sum())

    # message passing with the dgl.function builtins (gatedgcn_layer.py:48-53) or user-defined functions
    # (pna_layer.py:37-55,65-68: edges.src / edges.dst / edges.data, nodes.mailbox with DGL's degree bucketing)
    def apply_edges(self, func, edges=None):
        if callable(func):
            if edges is not None:   # the callers here pass every edge id (transformer.py:160-178 with full_graph=False)
                assert torch.equal(torch.as_tensor(edges), torch.arange(self.src.numel()))
            self.edata.update(func(_EdgeBatch(self)))
            return
        kind, u, v, out = func
        assert kind == "u_add_v"
        self.edata[out] = self.ndata[u].index_select(0, self.src) + self.ndata[v].index_select(0, self.dst)

    def send_and_recv(self, edges, message_func, reduce_func):
        """transformer.py:182-184 sends along g.edges() = every edge: identical to update_all."""
        self.update_all(message_func, reduce_func)

    def update_all(self, message_func, reduce_func):
        if callable(reduce_func):
            self._update_all_udf(message_func, reduce_func)
            return
        kind, a, b, m_name = message_func
        if kind in ("u_mul_e", "src_mul_edge"):
            m = self.ndata[a].index_select(0, self.src) * self.edata[b]
        elif kind in ("copy_e", "copy_edge"):
            m = self.edata[a]
        else:
            raise NotImplementedError(kind)
        rkind, r_in, r_out = reduce_func
        assert rkind == "sum" and r_in == m_name
        self.ndata[r_out] = torch.zeros(self.num_nodes(), *m.shape[1:], dtype=m.dtype).index_add(0, self.dst, m)


class _EdgeBatch:
    def __init__(self, g):
        self.src = _Gather(g.ndata, g.src)
        self.dst = _Gather(g.ndata, g.dst)
        self.data = g.edata


class _Gather:
    def __init__(self, store, index):
        self.store, self.index = store, index

    def __getitem__(self, key):
        return self.store[key].index_select(0, self.index)


class _NodeBatch:
    def __init__(self, mailbox):
        self.mailbox = mailbox


def _update_all_udf(self, message_func, reduce_func):
    """DGL semantics for a user-defined reduce: nodes are bucketed by in-degree D > 0 and the reduce function sees a
    mailbox [n_D, D, ...] whose messages are in edge-id order; nodes without incoming edges keep zeros."""
    if callable(message_func):
        msgs = message_func(_EdgeBatch(self))
    else:
        kind, a, _, m_name = message_func
        assert kind == "copy_u"
        msgs = {m_name: self.ndata[a].index_select(0, self.src)}
    N = self.num_nodes()
    order = torch.sort(self.dst, stable=True).indices            # edges grouped by destination, edge-id order inside
    deg = torch.bincount(self.dst, minlength=N)
    start = torch.cumsum(deg, 0) - deg
    out = {}
    for D in torch.unique(deg).tolist():
        if D == 0:
            continue
        nodes = torch.nonzero(deg == D).flatten()
        eids = order[(start[nodes].unsqueeze(1) + torch.arange(D).unsqueeze(0)).flatten()]
        mailbox = {k: v.index_select(0, eids).reshape(nodes.numel(), D, *v.shape[1:]) for k, v in msgs.items()}
        res = reduce_func(_NodeBatch(mailbox))
        for k, v in res.items():
            if k not in out:
                out[k] = torch.zeros(N, *v.shape[1:], dtype=v.dtype)
            out[k] = out[k].index_copy(0, nodes, v)
    self.ndata.update(out)


BatchedGraph._update_all_udf = _update_all_udf


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


def broadcast_nodes(g, x):
    """dgl.broadcast_nodes: repeat a per-graph row for every node of that graph (train_ZINC_graph_regression.py:41)."""
    return x.repeat_interleave(g.batch_num_nodes(), 0)

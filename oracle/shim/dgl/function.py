"""TEST INFRASTRUCTURE ONLY — the four dgl.function builtins GraphPrediction/layers/gatedgcn_layer.py:48-53 uses,
as plain descriptors interpreted by BatchedGraph.apply_edges / update_all (dgl/__init__.py).  Published semantics:
u_add_v(u, v, out): edata[out] = ndata[u][src] + ndata[v][dst];  u_mul_e(u, e, out): message = ndata[u][src] * edata[e];
copy_e(e, out): message = edata[e];  sum(msg, out): ndata[out][v] = sum of the messages of v's incoming edges."""


def u_add_v(u, v, out):
    return ("u_add_v", u, v, out)


def u_mul_e(u, e, out):
    return ("u_mul_e", u, e, out)


def copy_e(e, out):
    return ("copy_e", e, None, out)


def src_mul_edge(u, e, out):   # pre-0.5 name of u_mul_e (transformer.py:183)
    return ("src_mul_edge", u, e, out)


def copy_edge(e, out):         # pre-0.5 name of copy_e (transformer.py:184)
    return ("copy_edge", e, None, out)


def copy_u(u, out):
    return ("copy_u", u, None, out)


def sum(msg, out):  # noqa: A001 (dgl's own name)
    return ("sum", msg, out)

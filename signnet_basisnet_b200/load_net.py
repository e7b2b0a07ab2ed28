"""`gnn_model(MODEL_NAME, net_params)` with the names GraphPrediction/nets/ZINC_graph_regression/load_net.py:26-36 accepts,
for the predictors built on the B200 path: GIN and GatedGCN, PNA, Transformer (section 8f
rank 4).  GAT - dgl's own GATConv, a third-party op whose source is not part of the reference - is not built: asking for
it raises instead of silently running something else."""
from __future__ import annotations

import importlib

# model name -> (module of this package, class); imported on demand so that a predictor's import cost is paid only if used
_REGISTRY = {
    "GIN": ("gin_net", "GINNet"),
    "GatedGCN": ("gatedgcn_net", "GatedGCNNet"),
    "PNA": ("pna_net", "PNANet"),
    "Transformer": ("graph_transformer_net", "TransformerNet"),
}
_NOT_BUILT = {"GAT"}


def model_class(name: str):
    if name in _NOT_BUILT:
        raise NotImplementedError(f"{name} is not built on the B200 path (SURVEY section 8f rank 4, still open)")
    module, cls = _REGISTRY[name]   # KeyError for unknown names, like the reference's dict lookup
    return getattr(importlib.import_module(f"{__package__}.{module}"), cls)


def gnn_model(MODEL_NAME, net_params):
    return model_class(MODEL_NAME)(net_params)

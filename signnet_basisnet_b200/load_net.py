"""`gnn_model(MODEL_NAME, net_params)` as in GraphPrediction/nets/ZINC_graph_regression/load_net.py:11-36, for the
predictors built on the B200 path (GIN: SURVEY section 8 row a13; GatedGCN, PNA, Transformer: section 8f rank 4).  GAT
is not built: asking for them raises instead of silently running something else."""
from __future__ import annotations

from .gatedgcn_net import GatedGCNNet
from .gin_net import GINNet
from .graph_transformer_net import TransformerNet
from .pna_net import PNANet


def GatedGCN(net_params):
    return GatedGCNNet(net_params)


def GIN(net_params):
    return GINNet(net_params)


def PNA(net_params):
    return PNANet(net_params)


def Transformer(net_params):
    return TransformerNet(net_params)


def gnn_model(MODEL_NAME, net_params):
    models = {"GatedGCN": GatedGCN, "GIN": GIN, "PNA": PNA, "Transformer": Transformer}
    if MODEL_NAME in ("GAT",):
        raise NotImplementedError(f"{MODEL_NAME} is not built on the B200 path (SURVEY section 8f rank 4, still open)")
    return models[MODEL_NAME](net_params)   # KeyError for unknown names, like the reference

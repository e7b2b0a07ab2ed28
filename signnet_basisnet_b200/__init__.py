"""signnet_basisnet_b200 — B200-native (sm_100a) SignNet / BasisNet hot path behind the reference's module API.

PyG trees:           sign_net.SignNetGNN, sign_net.SignNet, model.GNN
DGL tree:            deepsigns.get_sign_inv_net (GINDeepSigns / MaskedGINDeepSigns), gin_net.GINNet
LearningFilters:     basisnet.SignPlus, basisnet.EqDeepSetsEncoder, basisnet.IGN2to1, basisnet.IGNBasisInv
Data path:           ops.to_dense_list_EVD, ops.laplacian_evd, ops.lap_positional_encoding
Multi-GPU:           ddp.shard_batch, ddp.FlatGradAllReduce

Everything computes through lib/libsignnet_b200.so (C ABI: include/signnet_b200.h); there is no CPU fallback.
Submodules are imported lazily so that `import signnet_basisnet_b200` needs neither a GPU nor the built library.
"""
__all__ = ["basisnet", "ddp", "deepsigns", "functional", "gin_net", "layout", "model", "ops", "phi", "sign_net", "synth",
           "transformer"]

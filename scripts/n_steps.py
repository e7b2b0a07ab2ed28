"""Run n plain steps of the bench model (for ncu launch lists):  python scripts/n_steps.py B n"""
import sys

import torch

sys.path.insert(0, ".")
import bench
from signnet_basisnet_b200.sign_net import SignNetGNN

B, n = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.manual_seed(0)
CFG = bench.CFG
model = SignNetGNN(None, None, CFG["n_hid"], CFG["n_out"], CFG["nl_signnet"], CFG["nl_gnn"], flavour=CFG["flavour"]).to(dev).train()
data = bench.make_batch(B, seed=1000).to(dev)
for i in range(n):
    for p in model.parameters():
        p.grad = None
    data.__dict__.pop("_b200_graph_index", None)
    out = model(data)
    (out - data.y).abs().mean().backward()
    torch.cuda.synchronize()
    print("step", i, flush=True)

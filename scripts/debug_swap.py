import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import restate
from signnet_basisnet_b200 import phi as phimod
from signnet_basisnet_b200.layout import GraphIndex, pad4
from signnet_basisnet_b200.sign_net import GNN3d, build_phi_input
from signnet_basisnet_b200.synth import synth_batch
from helpers import slot_row_index, dense_to_rows
DEV = "cuda"
nhid, nl = int(sys.argv[1]), 2
torch.manual_seed(0)
d = synth_batch(16, "zinc", seed=11)
phi = GNN3d(1, nhid, nl).to(DEV).train()
with torch.no_grad():
    for n_, p in phi.named_parameters():
        if n_.endswith("eps"): p.uniform_(-0.3, 0.3)
sd64 = {k: (v.detach().cpu().clone().double() if v.is_floating_point() else v.detach().cpu().clone()) for k, v in phi.state_dict().items()}
for k_, v in sd64.items():
    if v.is_floating_point() and "running_" not in k_: v.requires_grad_(True)
_, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
k = eigV.shape[1]; mask = restate.slot_mask(d.batch, k)
w = torch.randn(eigV.shape[0], k, nhid, generator=torch.Generator().manual_seed(1)) * mask.unsqueeze(-1)
ref = restate.phi_pm(eigV.double(), d.edge_index, mask, sd64, "", nl, True)
(ref * w.double()).sum().backward()
gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
sl = gi.slots_all(pad4(nhid))
idx = slot_row_index(d.batch, k, True)
w_rows = dense_to_rows(w, idx, pad4(nhid)).to(DEV)
res = []
for sgn in (1.0, -1.0):
    for p in phi.parameters(): p.grad = None
    x0 = build_phi_input(gi, sl, (sgn * d.eigen_vectors).to(DEV))
    xr, _ = phi.forward_rows(x0, gi, k, True)
    (xr * w_rows.unsqueeze(0)).sum().backward()
    res.append({n_: p.grad.cpu().double() for n_, p in phi.named_parameters() if p.grad is not None})
for n_ in res[0]:
    r = sd64[n_].grad
    s = max(r.abs().max().item(), 1e-30)
    print(f"{n_:40s} |ref| {s:10.3e}  err(+,-) {(res[0][n_]-r).abs().max().item()/s:9.2e}  err(-,+) {(res[1][n_]-r).abs().max().item()/s:9.2e}")

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
o_bnb = phimod.bn_backward
calls = []
def spy(gout, y, a, c, mr, gamma, ld, R, G, C, relu, training, dz_out):
    gin = gout.clone()
    dg, db = o_bnb(gout, y, a, c, mr, gamma, ld, R, G, C, relu, training, dz_out)
    z = a[:, None, :] * y[..., :C] + c[:, None, :]
    m = (z > 0)
    dz = gin[..., :C].double() * m
    s1 = dz.sum((0, 1)); yhat = (y[..., :C].double() - mr[0][:, None, :]) * mr[1][:, None, :]
    s2 = (dz * yhat).sum((0, 1))
    calls.append(dict(db=db.clone(), dg=dg.clone(), s1=s1, s2=s2, nz=(z == 0).sum().item(), near=(z.abs() < 1e-6).sum().item(), tot=z.numel()))
    return dg, db
phimod.bn_backward = spy
_, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
k = eigV.shape[1]; mask = restate.slot_mask(d.batch, k)
w = torch.randn(eigV.shape[0], k, nhid, generator=torch.Generator().manual_seed(1)) * mask.unsqueeze(-1)
gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
sl = gi.slots_all(pad4(nhid))
idx = slot_row_index(d.batch, k, True)
w_rows = dense_to_rows(w, idx, pad4(nhid)).to(DEV)
x0 = build_phi_input(gi, sl, d.eigen_vectors.to(DEV))
xr, _ = phi.forward_rows(x0, gi, k, True)
(xr * w_rows.unsqueeze(0)).sum().backward()
for i, c_ in enumerate(calls):
    e1 = ((c_["db"].double() - c_["s1"]).abs().max() / c_["s1"].abs().max()).item()
    e2 = ((c_["dg"].double() - c_["s2"]).abs().max() / c_["s2"].abs().max()).item()
    print(f"bn_backward call {i}: kernel-vs-torch dbeta {e1:.2e} dgamma {e2:.2e}; exact zeros in z {c_['nz']} near {c_['near']} of {c_['tot']}")

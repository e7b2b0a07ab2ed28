import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import restate
from signnet_basisnet_b200.layout import GraphIndex, pad4
from signnet_basisnet_b200.sign_net import GNN3d, build_phi_input
from signnet_basisnet_b200.synth import synth_batch
from helpers import slot_row_index, rows_to_dense
DEV = "cuda"
nhid, nl = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(0)
d = synth_batch(16, "zinc", seed=11)
phi = GNN3d(1, nhid, nl).to(DEV).train()
with torch.no_grad():
    for n_, p in phi.named_parameters():
        if n_.endswith("eps"): p.uniform_(-0.3, 0.3)
sd64 = {k: (v.detach().cpu().clone().double() if v.is_floating_point() else v.detach().cpu().clone()) for k, v in phi.state_dict().items()}
_, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
k = eigV.shape[1]; mask = restate.slot_mask(d.batch, k)
x = eigV.double().unsqueeze(-1)
refs = [restate.gnn3d(x, d.edge_index, mask, sd64, "", nl, True), restate.gnn3d(-x, d.edge_index, mask, sd64, "", nl, True)]
gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
sl = gi.slots_all(pad4(nhid))
idx = slot_row_index(d.batch, k, True)
x0 = build_phi_input(gi, sl, d.eigen_vectors.to(DEV))
with torch.no_grad():
    xr, _ = phi.forward_rows(x0, gi, k, True)
for s in (0, 1):
    got = rows_to_dense(xr[s].cpu(), idx, nhid).double()
    print(f"nhid {nhid} L {nl} sign {s}: fwd rel err {((got - refs[s]).abs().max() / refs[s].abs().max()).item():.3e}   vs other sign {((got - refs[1-s]).abs().max() / refs[s].abs().max()).item():.3e}")
print("sum err", ((rows_to_dense(xr[0].cpu(), idx, nhid).double() + rows_to_dense(xr[1].cpu(), idx, nhid).double() - refs[0] - refs[1]).abs().max() / (refs[0]+refs[1]).abs().max()).item())

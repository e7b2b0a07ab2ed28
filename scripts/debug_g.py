import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import restate
from signnet_basisnet_b200 import phi as phimod
from signnet_basisnet_b200.layout import GraphIndex, pad4
from signnet_basisnet_b200.sign_net import GNN3d, build_phi_input
from signnet_basisnet_b200.synth import synth_batch
from helpers import slot_row_index, dense_to_rows, rows_to_dense
DEV = "cuda"
shape, B, flavour, nhid, nl = "zinc", 16, "alchemy", 128, 2
torch.manual_seed(0)
d = synth_batch(B, shape, seed=11)
phi = GNN3d(1, nhid, nl, flavour=flavour).to(DEV).train()
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
with torch.no_grad():
    for n_, p in phi.named_parameters():
        if n_.endswith("bn.weight") and mode in ("all", "gamma"): p.uniform_(0.5, 1.5)
        elif n_.endswith("bn.bias") and mode in ("all", "beta"): p.uniform_(-0.3, 0.3)
        elif n_.endswith("eps") and mode in ("all", "eps"): p.uniform_(-0.3, 0.3)
rec = {}
o_agg, o_bnb, o_lin = phimod.gin_agg, phimod.bn_backward, phimod.linear_fwd
def spy_agg(x, out, slots, S, ld, **kw):
    pre = out.clone() if kw.get("res") is not None else None
    o_agg(x, out, slots, S, ld, **kw)
    if kw.get("transpose"):
        rec.setdefault("agg_bwd", []).append((x.clone(), pre, out.clone(), ld))
def spy_bnb(gout, y, a, c, mr, gamma, ld, R, G, C, relu, training, dz_out):
    gin = gout.clone()
    r = o_bnb(gout, y, a, c, mr, gamma, ld, R, G, C, relu, training, dz_out)
    rec.setdefault("bnb", []).append((gin, y.clone(), dz_out.clone(), C))
    return r
phimod.gin_agg, phimod.bn_backward = spy_agg, spy_bnb
sd64 = {k: (v.detach().cpu().clone().double() if v.is_floating_point() else v.detach().cpu().clone()) for k, v in phi.state_dict().items()}
for k_, v in sd64.items():
    if v.is_floating_point() and "running_" not in k_: v.requires_grad_(True)
_, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
k = eigV.shape[1]; mask = restate.slot_mask(d.batch, k)
w = torch.randn(eigV.shape[0], k, nhid, generator=torch.Generator().manual_seed(1)) * mask.unsqueeze(-1)
# oracle with captured intermediates: re-implement loop to retain grads
caps = []
import torch.nn.functional as F
def gnn3d_cap(x, sign):
    x = x.transpose(0, 1); m = mask.transpose(0, 1); prev = 0; out = {}
    for l in range(nl):
        p = ""
        a = restate.gin_aggregate(x, d.edge_index, sd64[f"convs.{l}.layer.eps"]); a.retain_grad() if a.requires_grad else None
        h = F.linear(a, sd64[f"convs.{l}.nn.layers.0.weight"]) * m.unsqueeze(-1); h.retain_grad()
        hn = F.relu(restate._masked_bn(h, m, sd64, f"convs.{l}.nn.norms.0.", True))
        y = F.linear(hn, sd64[f"convs.{l}.nn.layers.1.weight"], sd64.get(f"convs.{l}.nn.layers.1.bias")) * m.unsqueeze(-1); y.retain_grad()
        z = F.relu(restate._masked_bn(y, m, sd64, f"norms.{l}.", True))
        x = z + prev; x.retain_grad(); prev = x
        out[l] = dict(A=a, H=h, Y=y, X=x)
    caps.append(out)
    return x.transpose(0, 1)
xin = eigV.double().unsqueeze(-1)
ref = gnn3d_cap(xin, 0) + gnn3d_cap(-xin, 1)
(ref * w.double()).sum().backward()
gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
sl = gi.slots_all(pad4(nhid))
x0 = build_phi_input(gi, sl, d.eigen_vectors.to(DEV))
xr, sl = phi.forward_rows(x0, gi, k, True)
idx = slot_row_index(d.batch, k, True)
w_rows = dense_to_rows(w, idx, pad4(nhid)).to(DEV)
(xr * w_rows.unsqueeze(0)).sum().backward()
def cmp(name, got_rows, ref_dense):  # got [R, ld], ref [k, N, C]
    C = ref_dense.shape[-1]
    r = dense_to_rows(ref_dense.transpose(0, 1), idx, C)
    g = got_rows[:, :C].double().cpu()
    print(f"  {name}: rel err {((g - r).abs().max() / r.abs().max()).item():.3e}")
# bnb calls order: layer1 outer (G->dY), layer1 inner (dP->dH), layer0 outer, layer0 inner
names = ["L1 outer", "L1 inner", "L0 outer", "L0 inner"]
for i, (gin, y, dz, C) in enumerate(rec["bnb"]):
    l = 1 - i // 2
    for s in (0, 1):
        c = caps[s][l]
        if i % 2 == 0:
            cmp(f"{names[i]} sign{s} gout(dX_{l+1})", gin[s], c["X"].grad)
            cmp(f"{names[i]} sign{s} dY", dz[s], c["Y"].grad)
        else:
            cmp(f"{names[i]} sign{s} dH", dz[s], c["H"].grad)
for i, (x, pre, out, ld) in enumerate(rec["agg_bwd"]):
    print("agg_bwd call", i, "ld", ld)
    if ld > 1:
        for s in (0, 1):
            cmp(f"  sign{s} dA (input)", x[s], caps[s][1]["A"].grad)
            cmp(f"  sign{s} G out (dX_1)", out[s], caps[s][0]["X"].grad)

import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import restate
from signnet_basisnet_b200.sign_net import SignNetGNN
from signnet_basisnet_b200.synth import synth_batch
DEV = "cuda"
shape, B, nhid = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(7)
d = synth_batch(B, shape, seed=31)
nf, ef = (6, 4) if shape == "alchemy" else (None, None)
if shape == "zinc": d.x, d.edge_attr = d.x % 6, d.edge_attr % 6
model = SignNetGNN(nf, ef, n_hid=nhid, n_out=3, nl_signnet=2, nl_gnn=2).to(DEV).train()
for lyr in model.sign_net.rho.transformer_layers: lyr.slf_attn.attention.dropout.p = 0.0
sd64 = {k: (v.detach().cpu().clone().double() if v.is_floating_point() else v.detach().cpu().clone()) for k, v in model.state_dict().items()}
for k, v in sd64.items():
    if v.is_floating_point() and "running_" not in k: v.requires_grad_(True)
d64 = d.to("cpu")
for k in ("x", "edge_attr", "eigen_values", "eigen_vectors"):
    v = getattr(d64, k)
    if v.is_floating_point(): setattr(d64, k, v.double())
ref = restate.sign_net_gnn(d64, sd64, 2, 2); ref.abs().sum().backward()
out = model(d.to(DEV)); out.abs().sum().backward()
print("out err", ((out.cpu().double() - ref).abs().max() / ref.abs().max()).item())
gmax = max(float(v.grad.abs().max()) for v in sd64.values() if v.requires_grad and v.grad is not None)
for n_, p in model.named_parameters():
    r = sd64[n_].grad
    if p.grad is None or r is None:
        if (p.grad is None) != (r is None): print("NONE MISMATCH", n_, p.grad is None, r is None)
        continue
    e = (p.grad.cpu().double() - r).abs().max().item() / max(r.abs().max().item(), 0.1 * gmax)
    if e > 5e-6: print(f"{n_:60s} ref {r.abs().max().item():9.3e} err {e:9.2e}")
for n_, p in model.named_parameters():
    if "eigen_encoder" in n_ and p.numel() <= 2:
        print(n_, "got", p.grad.flatten().tolist() if p.grad is not None else None, "ref", sd64[n_].grad.flatten().tolist() if sd64[n_].grad is not None else None)
print("gmax", gmax)

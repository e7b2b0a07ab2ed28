import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import restate
from signnet_basisnet_b200 import functional as Fn
from signnet_basisnet_b200.sign_net import SignNetGNN
from signnet_basisnet_b200.synth import synth_batch
DEV = "cuda"
torch.manual_seed(7)
d = synth_batch(5, "alchemy", seed=31)
model = SignNetGNN(6, 4, n_hid=16, n_out=3, nl_signnet=2, nl_gnn=2).to(DEV).train()
for lyr in model.sign_net.rho.transformer_layers: lyr.slf_attn.attention.dropout.p = 0.0
o_bnb = Fn.bn_backward
def spy(gout, y, a, c, mr, gamma, ld, R, G, C, relu, training, dz_out):
    gin = gout.clone()
    dg, db = o_bnb(gout, y, a, c, mr, gamma, ld, R, G, C, relu, training, dz_out)
    if C == 1:
        z = a[:, None, :] * y.view(G, R, ld)[..., :C] + c[:, None, :]
        print("C=1 BN backward: R", R, "z==0:", (z == 0).sum().item(), "|z|<1e-6:", (z.abs() < 1e-6).sum().item(), "z>0:", (z > 0).sum().item(),
              "db", db.tolist(), "sum gin*(z>0)", (gin.view(G, R, ld)[..., :C] * (z > 0)).sum().item(), "sum gin*(z>=0)", (gin.view(G, R, ld)[..., :C] * (z >= 0)).sum().item())
        print("   y unique (first 12):", torch.unique(y.view(R, ld)[:, 0])[:12].tolist(), "a", a.tolist(), "c", c.tolist(), "mean", mr[0].tolist())
    return dg, db
Fn.bn_backward = spy
out = model(d.to(DEV)); out.abs().sum().backward()

"""On-GPU probe of the GatedGCN + SignNet(masked GIN) predictor at the shape of
GraphPrediction/configs/gatedgcn/GatedGCN_ZINC_LapPE_signinv_GIN_mask.json (hidden 67, 16 GatedGCN layers, k = 37,
8 phi layers, pe_aggregate concat): fwd+bwd ms/step, graphs/s, per-entry-point CUDA-event breakdown and the algorithmic
bandwidth of the edge-gated aggregate (sb_gated_agg_fwd reads 4 node rows + 1 edge row per edge-feature, writes 1 edge
row + 3 node rows:  bytes = 4 * ld * (2 E + 2 * 2 E [Dh, Bh gathers] + 6 N) per launch, counted below).
    python scripts/gatedgcn_probe.py [B]"""
import sys

import torch

sys.path.insert(0, ".")
from signnet_basisnet_b200 import _lib
from signnet_basisnet_b200.gatedgcn_net import GatedGCNNet, handle_lap
from signnet_basisnet_b200.layout import pad4
from signnet_basisnet_b200.synth import synth_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = "cuda"
prm = dict(num_atom_type=28, num_bond_type=4, hidden_dim=67, out_dim=67, in_feat_dropout=0.0, dropout=0.0, L=16,
           readout="mean", batch_norm=True, residual=True, edge_feat=True, device=dev, pe_init="lap_pe",
           lap_method="sign_inv", lap_lspe=False, use_lapeig_loss=False, lambda_loss=1.0, alpha_loss=1e-4, pos_enc_dim=37,
           sign_inv_net="masked_gin", phi_out_dim=67, sign_inv_layers=8, sign_inv_activation="relu", pe_aggregate="concat")
torch.manual_seed(0)
net = GatedGCNNet(prm).to(dev).train()
d = synth_batch(B, "zinc", seed=0, k_dgl=prm["pos_enc_dim"]).to(dev)
y = torch.randn(B, 1, device=dev)


class G:
    def edges(self):
        return d.edge_index[0], d.edge_index[1]

    def batch_num_nodes(self):
        return torch.as_tensor(d.num_nodes_per_graph)


g = G()


def step():
    for p in net.parameters():
        p.grad = None
    pe = handle_lap(net, d.pos_enc, g, dev)
    out, _ = net(g, d.x[:, 0], pe, d.edge_attr.reshape(-1), None)
    net.loss(out, y).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
N, E = int(d.batch.numel()), int(d.edge_index.shape[1])
print(f"GatedGCN + SignNet fwd+bwd: B={B} N={N} E={E}: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} graphs/s", flush=True)
_lib.profile_start()
step()
prof = _lib.profile_stop()
tot = sum(t for _, t in prof.values())
for tag, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{tag:44s} calls {c:4d}  {t:9.3f} ms  {100 * t / tot:5.1f}%  avg {t / c * 1e3:9.1f} us")
ld = pad4(prm["hidden_dim"])
if "sb_gated_agg_fwd" in prof:
    c, t = prof["sb_gated_agg_fwd"]
    byt = 4 * ld * (2 * E + 2 * E + 6 * N)
    print(f"gated aggregate fwd: {byt / (t / c * 1e-3) / 1e9:.0f} GB/s algorithmic ({byt / 1e6:.1f} MB per launch)")
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)

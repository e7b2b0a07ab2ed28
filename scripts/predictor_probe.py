"""On-GPU probe of the three DGL predictors of SURVEY §8f rank 4 with their SignNet (masked GIN) positional encoder, at the
shapes the reference ships (GraphPrediction/configs/{gatedgcn,pna,transformer}/*_signinv_GIN_mask*.json: batch 128,
k = 37, 8 phi layers): fwd+bwd ms/step, graphs/s, per-entry-point CUDA-event breakdown and the algorithmic HBM bandwidth
of each predictor's aggregate kernel against  2*4*d*N + 16*E (+ 4*d*E per edge-feature tensor read or written).
    python scripts/predictor_probe.py [gatedgcn|pna|transformer] [B]"""
import sys

import torch

sys.path.insert(0, ".")
from signnet_basisnet_b200 import _lib
from signnet_basisnet_b200.gatedgcn_net import GatedGCNNet, handle_lap
from signnet_basisnet_b200.layout import pad4
from signnet_basisnet_b200.synth import synth_batch

which = sys.argv[1] if len(sys.argv) > 1 else "gatedgcn"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = "cuda"
common = dict(num_atom_type=28, num_bond_type=4, in_feat_dropout=0.0, dropout=0.0, batch_norm=True, residual=True,
              edge_feat=True, device=dev, pe_init="lap_pe", lap_method="sign_inv", lap_lspe=False, use_lapeig_loss=False,
              alpha_loss=1e-4, pos_enc_dim=37, sign_inv_net="masked_gin", sign_inv_layers=8, sign_inv_activation="relu",
              pe_aggregate="concat")
if which == "gatedgcn":
    prm = dict(common, hidden_dim=67, out_dim=67, L=16, readout="mean", lambda_loss=1.0, phi_out_dim=67)
    net = GatedGCNNet(prm)
    agg, n_edge_tensors = "sb_gated_agg_fwd", 2      # Ce read, e_out written
elif which == "pna":
    from signnet_basisnet_b200.pna_net import PNANet
    prm = dict(common, hidden_dim=70, out_dim=70, L=16, readout="sum", graph_norm=True, aggregators="mean max min std",
               scalers="identity amplification attenuation", avg_d={"log": 1.1}, towers=5, divide_input_first=True,
               divide_input_last=True, gru=False, edge_dim=40, pretrans_layers=1, posttrans_layers=1, lambda_loss=1000,
               phi_out_dim=70)
    net = PNANet(prm)
    agg, n_edge_tensors = "sb_pna_agg_fwd", 1        # Q read
else:
    from signnet_basisnet_b200.graph_transformer_net import TransformerNet
    prm = dict(common, hidden_dim=56, out_dim=56, n_heads=8, full_graph=False, L=10, readout="sum", layer_norm=True,
               lambda_loss=1, phi_out_dim=16)
    net = TransformerNet(prm)
    agg, n_edge_tensors = "sb_edge_attention_fwd", 1  # E read
torch.manual_seed(0)
net = net.to(dev).train()
d = synth_batch(B, "zinc", seed=0, k_dgl=prm["pos_enc_dim"]).to(dev)
y = torch.randn(B, 1, device=dev)
snorm = (1.0 / torch.as_tensor(d.num_nodes_per_graph, dtype=torch.float32).sqrt()).repeat_interleave(
    torch.as_tensor(d.num_nodes_per_graph)).unsqueeze(1).to(dev)


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
    out, _ = net(g, d.x[:, 0], pe, d.edge_attr.reshape(-1), snorm if which == "pna" else None)
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
print(f"{which} + SignNet(masked GIN, k=37, 8 layers) fwd+bwd: B={B} N={N} E={E}: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} graphs/s",
      flush=True)
_lib.profile_start()
step()
prof = _lib.profile_stop()
tot = sum(t for _, t in prof.values())
for tag, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{tag:46s} calls {c:4d}  {t:9.3f} ms  {100 * t / tot:5.1f}%  avg {t / c * 1e3:9.1f} us")
ld = pad4(prm["hidden_dim"])
for tag in (agg, agg.replace("_fwd", "_bwd")):
    if tag in prof:
        c, t = prof[tag]
        byt = 2 * 4 * ld * N + 16 * E + n_edge_tensors * 4 * ld * E
        gbs = byt / (t / c * 1e-3) / 1e9
        print(f"{tag}: {t / c * 1e3:.1f} us per launch, {byt / 1e6:.2f} MB algorithmic (2*4*d*N + 16*E + {n_edge_tensors}*4*d*E) "
              f"-> {gbs:.0f} GB/s = {gbs / 6543.4:.3f} of the measured HBM peak")
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)

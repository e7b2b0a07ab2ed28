"""Quick on-GPU probe: phi fwd+bwd at the cfg-4 shape with a per-entry-point CUDA-event breakdown."""
import sys
import time

import torch

sys.path.insert(0, ".")
from signnet_basisnet_b200 import _lib
from signnet_basisnet_b200.layout import GraphIndex, pad4
from signnet_basisnet_b200.sign_net import GNN3d, build_phi_input
from signnet_basisnet_b200.synth import synth_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d_hid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
L = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = "cuda"
t0 = time.time()
d = synth_batch(B, "zinc", seed=0).to(dev)
print("synth", time.time() - t0, flush=True)
phi = GNN3d(1, d_hid, L).to(dev).train()
gi = GraphIndex(d.edge_index, d.batch, d.num_graphs)
sl = gi.slots_all(pad4(d_hid))
print("N", gi.N, "E", gi.E, "R", sl.R, "nmax", sl.nmax, flush=True)
x0 = build_phi_input(gi, sl, d.eigen_vectors)


def step():
    for p in phi.parameters():
        p.grad = None
    xr, _ = phi.forward_rows(x0, gi, sl.k, True)
    xr.sum().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"phi fwd+bwd: {ms:.2f} ms/step  -> {B / ms * 1e3:.0f} graphs/s", flush=True)
_lib.profile_start()
step()
prof = _lib.profile_stop()
tot = sum(t for _, t in prof.values())
for tag, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"{tag:40s} calls {c:4d}  {t:9.3f} ms  {100 * t / tot:5.1f}%  avg {t / c * 1e3:9.1f} us")
R, ld = sl.R, pad4(d_hid)
agg = [v for k_, v in prof.items() if k_.startswith(f"sb_gin_agg[ld={ld}")]
for (c, t) in agg:
    byt = 2 * 4 * ld * 2 * R + 16 * gi.E
    print(f"agg: {byt / (t / c * 1e-3) / 1e9:.0f} GB/s algorithmic")
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)

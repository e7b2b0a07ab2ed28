"""Host-side profile of cfg 2 (Alchemy SignNetGNN, B=128, d=64, 8 phi layers, 16 GINE layers, 4 rho layers)."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from signnet_basisnet_b200 import _lib
from signnet_basisnet_b200.sign_net import SignNetGNN
from signnet_basisnet_b200.synth import synth_batch

torch.manual_seed(0)
d = synth_batch(128, "alchemy", seed=7).to("cuda")
model = SignNetGNN(6, 4, 64, 12, 8, 16).to("cuda").train()
params = list(model.parameters())


def step():
    for p in params:
        p.grad = None
    d.__dict__.pop("_b200_graph_index", None)
    model(d).abs().mean().backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
c0 = _lib.launch_count
t0 = time.perf_counter()
step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"cfg2: cpu issue {1e3 * (t1 - t0):.2f} ms, wall {1e3 * (t2 - t0):.2f} ms, C-ABI calls {_lib.launch_count - c0}")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr, stream=sys.stdout).sort_stats("tottime").print_stats(22)

"""Host-side cost of one step: CPU wall clock of forward / backward (GPU idle-waiting excluded by construction: nothing
synchronises inside), GPU time of the same step, and a cProfile of the Python side.  `python scripts/host_probe.py B`."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import bench
from signnet_basisnet_b200 import _lib
from signnet_basisnet_b200.sign_net import SignNetGNN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
torch.manual_seed(0)
CFG = bench.CFG
model = SignNetGNN(None, None, CFG["n_hid"], CFG["n_out"], CFG["nl_signnet"], CFG["nl_gnn"], flavour=CFG["flavour"]).to(dev).train()
data = bench.make_batch(B, seed=1000).to(dev)
params = list(model.parameters())


def step(log=False):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for p in params:
        p.grad = None
    data.__dict__.pop("_b200_graph_index", None)
    out = model(data)
    t1 = time.perf_counter()
    loss = (out - data.y).abs().mean()
    loss.backward()
    t2 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if log:
        print(f"B={B}: cpu forward {1e3*(t1-t0):.2f} ms, cpu backward {1e3*(t2-t1):.2f} ms, cpu issue total {1e3*(t2-t0):.2f} ms, "
              f"wall incl. sync {1e3*(t3-t0):.2f} ms, gpu {e0.elapsed_time(e1):.2f} ms")


for i in range(8):
    step(i >= 5)
c0 = _lib.launch_count
step()
print("C-ABI calls per step:", _lib.launch_count - c0)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
st = pstats.Stats(pr, stream=sys.stdout)
st.sort_stats("tottime").print_stats(28)

# free-running loop (what bench.py's `value` times): no synchronisation between steps
for sync_each in (False, True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        for p in params:
            p.grad = None
        data.__dict__.pop("_b200_graph_index", None)
        out = model(data)
        loss = (out - data.y).abs().mean()
        loss.backward()
        if sync_each:
            loss.item()
    torch.cuda.synchronize()
    print(f"20 steps, sync_each={sync_each}: {1e3 * (time.perf_counter() - t0) / 20:.2f} ms/step")

# pipelined bookkeeping (layout.prepare_batch on a side stream, one step ahead)
from signnet_basisnet_b200.layout import pad4, prepare_batch
side = torch.cuda.Stream()
ring = [type(data)(**data.__dict__), type(data)(**data.__dict__)]
for r in ring:
    r.__dict__.pop("_b200_graph_index", None)
LD = pad4(CFG["n_hid"])
for rep in range(2):
    torch.cuda.synchronize()
    prepare_batch(ring[0], LD)
    t0 = time.perf_counter()
    tp = 0.0
    for i in range(20):
        a, b = ring[i & 1], ring[(i + 1) & 1]
        for p in params:
            p.grad = None
        t1 = time.perf_counter()
        prepare_batch(b, LD, stream=side)
        tp += time.perf_counter() - t1
        out = model(a)
        a.__dict__.pop("_b200_graph_index", None)
        loss = (out - a.y).abs().mean()
        loss.backward()
    torch.cuda.synchronize()
    print(f"20 steps, pipelined bookkeeping: {1e3 * (time.perf_counter() - t0) / 20:.2f} ms/step (prepare_batch cpu {1e3 * tp / 20:.2f} ms/step)")

"""Where does the end-to-end step spend its time?  CPU wall-clock vs GPU event timeline of one bench.py e2e step."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench
from signnet_basisnet_b200.sign_net import SignNetGNN

dev = torch.device("cuda", 0)
torch.manual_seed(0)
CFG = bench.CFG
model = SignNetGNN(None, None, CFG["n_hid"], CFG["n_out"], CFG["nl_signnet"], CFG["nl_gnn"], flavour=CFG["flavour"]).to(dev).train()
host = bench.make_batch(1024, seed=1000).pin_memory()


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def step(log):
    t = [time.perf_counter()]
    e = [ev()]
    for p in model.parameters():
        p.grad = None
    data = host.to(dev, non_blocking=True)
    t.append(time.perf_counter()); e.append(ev())
    out = model(data)
    t.append(time.perf_counter()); e.append(ev())
    loss = (out - data.y).abs().mean()
    loss.backward()
    t.append(time.perf_counter()); e.append(ev())
    v = float(loss.item())
    t.append(time.perf_counter()); e.append(ev())
    torch.cuda.synchronize()
    if log:
        names = ["h2d", "forward", "backward", "item"]
        for i, n in enumerate(names):
            print(f"{n:9s} cpu {1e3 * (t[i + 1] - t[i]):7.2f} ms   gpu {e[i].elapsed_time(e[i + 1]):7.2f} ms")
        print(f"total     cpu {1e3 * (t[-1] - t[0]):7.2f} ms   gpu {e[0].elapsed_time(e[-1]):7.2f} ms")
    return v


for i in range(6):
    step(i >= 4)

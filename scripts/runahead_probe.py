"""How far ahead of the GPU does the issuing thread get?  Host time per step of 12 back-to-back steps, (a) with the
per-step bookkeeping rebuild on the side stream (bench.py's value loop), (b) reusing one prepared batch (no host sync at all)."""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from signnet_basisnet_b200.layout import pad4, prepare_batch
from signnet_basisnet_b200.sign_net import SignNetGNN
dev = torch.device("cuda", 0)
torch.manual_seed(0)
CFG = bench.CFG
model = SignNetGNN(None, None, CFG["n_hid"], CFG["n_out"], CFG["nl_signnet"], CFG["nl_gnn"], flavour=CFG["flavour"]).to(dev).train()
data = bench.make_batch(1024, seed=1000).to(dev)
params = list(model.parameters())
LD = pad4(CFG["n_hid"])
side = torch.cuda.Stream(device=dev, priority=-1)
def one(d):
    for p in params: p.grad = None
    out = model(d)
    (out - d.y).abs().mean().backward()
for mode in ("rebuild", "reuse"):
    twins = [type(data)(**data.__dict__), type(data)(**data.__dict__)]
    for t in twins: t.__dict__.pop("_b200_graph_index", None)
    prepare_batch(twins[0], LD); prepare_batch(twins[1], LD)
    for _ in range(3): one(twins[0])
    torch.cuda.synchronize()
    tt = []
    t_all = time.perf_counter()
    for i in range(12):
        t0 = time.perf_counter()
        a, b = twins[i & 1], twins[(i + 1) & 1]
        if mode == "rebuild":
            b.__dict__.pop("_b200_graph_index", None)
            prepare_batch(b, LD, stream=side)
        one(a)
        tt.append(1e3 * (time.perf_counter() - t0))
    t_issue = 1e3 * (time.perf_counter() - t_all)
    torch.cuda.synchronize()
    t_total = 1e3 * (time.perf_counter() - t_all)
    print(mode, "host ms per step:", " ".join(f"{t:.1f}" for t in tt), f"| issued in {t_issue:.0f} ms, done in {t_total:.0f} ms")

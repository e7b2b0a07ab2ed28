"""Throughput of the five BASELINE.json configs on one B200 (forward + backward of the model each config names, seeded
synthetic inputs of SURVEY §8d), next to the CPU oracle on the host cores for the same tensors.
    python scripts/configs_probe.py [--no-cpu]"""
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import restate  # noqa: E402  (timing baseline only)
from signnet_basisnet_b200.sign_net import SignNetGNN  # noqa: E402
from signnet_basisnet_b200.synth import synth_batch  # noqa: E402

DEV = "cuda"
CPU = "--no-cpu" not in sys.argv
torch.set_num_threads(torch.get_num_threads())


def gpu_time(step, n=20, warm=5):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def cpu_time(step, n=2):
    step()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    return (time.perf_counter() - t0) / n * 1e3


def leaf_sd(model):
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    return sd


def signnet_cfg(name, shape, B, nf, ef, nh, nout, L, G, flavour, k=None):
    torch.manual_seed(0)
    d = synth_batch(B, shape, seed=7)
    model = SignNetGNN(nf, ef, nh, nout, L, G, flavour=flavour).to(DEV).train()
    dd = d.to(DEV)
    eigV = None
    if k is not None:
        _, eigV = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
        eigV = eigV[:, :k].contiguous()
        eg = eigV.to(DEV)

    def step():
        for p in model.parameters():
            p.grad = None
        dd.__dict__.pop("_b200_graph_index", None)
        out = model(dd) if k is None else model(dd.x, dd.edge_index, eg, dd.batch, dd.edge_attr)
        out.abs().mean().backward()

    ms = gpu_time(step)
    line = f"{name}: B={B} N={d.batch.numel()} E={d.edge_index.shape[1]}  GPU {ms:8.2f} ms/step = {B / ms * 1e3:9.0f} graphs/s"
    if CPU:
        sd = leaf_sd(model)
        rho = 4 if flavour == "alchemy" else 1

        def cstep():
            for v in sd.values():
                if v.requires_grad:
                    v.grad = None
            if k is None:
                out = restate.sign_net_gnn(d, sd, L, G, nl_rho=rho, ignore_eigval=(flavour == "zinc"))
            else:
                pos = restate.sign_net(None, eigV, d.edge_index, d.batch, sd, "sign_net.", L, rho, True, True)
                out = restate.gnn_predictor(d.x, d.edge_index, d.edge_attr, d.batch, pos, sd, "gnn.", G, num_graphs=B)
            out.abs().mean().backward()

        cms = cpu_time(cstep)
        line += f" | CPU oracle ({torch.get_num_threads()} threads) {cms:9.1f} ms = {B / cms * 1e3:7.1f} graphs/s  ({cms / ms:.0f}x)"
    print(line, flush=True)


def cfg1():
    """LearningFilters sign_inv: phi = SignPlus(EqDeepSets(1,32,1,3)) on [k=8, n=200, 1], rho = EqDeepSets(16,10,32,3)."""
    from signnet_basisnet_b200.basisnet import EqDeepSetsEncoder, SignPlus

    torch.manual_seed(0)
    phi = SignPlus(EqDeepSetsEncoder(1, 32, 1, 3, use_bn=True)).to(DEV).train()
    rho = EqDeepSetsEncoder(16, 10, 32, 3, use_bn=True).to(DEV).train()
    V = torch.linalg.qr(torch.randn(200, 8))[0].t().contiguous().unsqueeze(-1).to(DEV)     # [k, n, 1]

    def step():
        for p in list(phi.parameters()) + list(rho.parameters()):
            p.grad = None
        z = phi(V)                                    # [k, n, 1]
        x = torch.cat([z.squeeze(-1).t(), V.squeeze(-1).t()], dim=1).contiguous()   # [n, 2k]
        rho(x).abs().mean().backward()

    ms = gpu_time(step)
    print(f"cfg1 LearningFilters single graph (n=200, k=8): GPU {ms:8.3f} ms/step (CPU-runnable parity config; launch-bound)", flush=True)


def cfg5():
    """BasisNet IGN-phi, k=16 lowest eigenvectors of a 25x40 grid grouped into eigenspaces (5-decimal rounding)."""
    from signnet_basisnet_b200.basisnet import IGNBasisInv, eigenspace_groups

    n1, n2 = 25, 40
    n = n1 * n2
    idx = torch.arange(n).view(n1, n2)
    src = torch.cat([idx[:, :-1].reshape(-1), idx[:-1, :].reshape(-1)])
    dst = torch.cat([idx[:, 1:].reshape(-1), idx[1:, :].reshape(-1)])
    A = torch.zeros(n, n, dtype=torch.float64)
    A[src, dst] = 1
    A = A + A.t()
    dg = A.sum(1)
    Lap = torch.eye(n, dtype=torch.float64) - A / (dg[:, None] * dg[None, :]).sqrt()
    ev, V = torch.linalg.eigh(Lap)
    ev, V = ev[:16].float(), V[:, :16].float().contiguous()
    groups = eigenspace_groups(ev)
    net = IGNBasisInv(sorted(groups), 1, hidden_channels=32).to(DEV).train()
    Vd = V.to(DEV)

    def step():
        for p in net.parameters():
            p.grad = None
        tot = 0
        for m, starts in groups.items():
            tot = tot + net.forward_factors(Vd, starts.to(DEV), m).abs().mean()
        tot.backward()

    ms = gpu_time(step)
    mult = {int(m): int(s.numel()) for m, s in groups.items()}
    print(f"cfg5 BasisNet IGN-phi (1000-node grid, k=16, eigenspaces by multiplicity {mult}): GPU {ms:8.3f} ms/step, from "
          f"eigenvector factors (projectors never materialised: 4*N*sum(mult) = {4 * n * 16 / 1e3:.0f} KB read vs "
          f"{4 * len(ev) * n * n / 1e6:.0f} MB of [b,1,N,N] projectors)", flush=True)


cfg1()
signnet_cfg("cfg2 Alchemy SignNet-GIN k=N_max(<=12) hidden=64", "alchemy", 128, 6, 4, 64, 12, 8, 16, "alchemy")
signnet_cfg("cfg3 ZINC GINESignNetPyG k=8 hidden=95", "zinc", 256, None, None, 95, 1, 4, 6, "zinc", k=8)
signnet_cfg("cfg4 ZINC SignNet-GIN k=37 hidden=128 (one GPU's share at 8 GPUs)", "zinc", 128, None, None, 128, 1, 8, 6, "zinc")
if "--full" in sys.argv:
    signnet_cfg("cfg4 ZINC SignNet-GIN k=37 hidden=128, B=1024", "zinc", 1024, None, None, 128, 1, 8, 6, "zinc")
cfg5()

"""On-GPU micro-benchmark of the phi aggregate (K1) at the cfg-4 shape: forward (x -> out) and backward mode
(out = res + agg^T(x), d eps dot), CUDA events over `iters` back-to-back launches on rotating buffers (> L2)."""
import sys

import torch

sys.path.insert(0, ".")
from signnet_basisnet_b200.layout import GraphIndex, pad4
from signnet_basisnet_b200.phi import gin_agg
from signnet_basisnet_b200.synth import synth_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d_hid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = "cuda"
d = synth_batch(B, "zinc", seed=1000).to(dev)
gi = GraphIndex(d.edge_index, d.batch, d.num_graphs)
ld = pad4(d_hid)
sl = gi.slots_all(ld)
R = sl.R
print(f"B {B} N {gi.N} E {gi.E} R {R} ld {ld} tile_rows {sl.tile_rows}", flush=True)
bufs = [torch.randn(2, R, ld, device=dev) for _ in range(4)]
eps = torch.zeros(1, device=dev)
deps = torch.zeros(1, dtype=torch.float64, device=dev)


def timed(fn):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def fwd(i):
    gin_agg(bufs[i % 2], bufs[2 + i % 2], sl, 2, ld, eps=eps)


def bwd(i):
    gin_agg(bufs[0], bufs[1], sl, 2, ld, eps=eps, res=bufs[1], dotx=bufs[2 + i % 2], dot_out=deps, transpose=True)


def fwd_generic(i):
    gin_agg(bufs[i % 2], bufs[2 + i % 2], sl, 2, ld, eps=eps, force_generic=True)


peak = 6543.4
for name, fn, nb in (("fwd", fwd, 2), ("bwd", bwd, 4), ("fwd_generic", fwd_generic, 2)):
    ms = timed(fn)
    byt = nb * 4 * ld * 2 * R + 16 * gi.E
    gbs = byt / (ms * 1e-3) / 1e9
    print(f"agg {name:12s} {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s algorithmic  {gbs / peak:.3f} of measured HBM peak", flush=True)

"""Timing of the fused aggregate -> Linear kernel against the two kernels it replaces, cfg 4 size (B graphs, ld = 128).
python scripts/fused_bench.py [B]"""
import sys

import torch

sys.path.insert(0, ".")
import bench
from signnet_basisnet_b200 import _lib
from signnet_basisnet_b200._lib import counted_call as call, ptr as p
from signnet_basisnet_b200.functional import linear_fwd
from signnet_basisnet_b200.layout import GraphIndex
from signnet_basisnet_b200.phi import gin_agg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = "cuda"
d = bench.make_batch(B, seed=1000)
gi = GraphIndex(d.edge_index.to(dev), d.batch.to(dev), d.num_graphs)
sl = gi.slots_all(128)
S, R = 2, sl.R
X = torch.randn(S, R, 128, device=dev)
W = (torch.randn(128, 128) / 128 ** 0.5).to(dev)
eps = torch.tensor([0.1], device=dev)
A, H = torch.empty_like(X), torch.empty_like(X)
st = torch.zeros(S, 2, 128, dtype=torch.float64, device=dev)
T = X.numel() * 4


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def two():
    gin_agg(X, A, sl, S, 128, eps=eps)
    linear_fwd(A, 128, W, 128, 1, None, H, 128, R, S, 128, 128, stats=st)


def fused():
    call("sb_gin_linear_fused_fwd", p(X), p(A), p(H), p(st), p(eps), p(W), 128, 1, 128, 128, 128, p(sl.unit_ptr),
         p(sl.unit_desc), p(gi.in_pack), p(gi.in_ptr), p(gi.in_src), R, gi.B, S, 128, sl.tile_rows, 0)


t_agg = timeit(lambda: gin_agg(X, A, sl, S, 128, eps=eps))
t_lin = timeit(lambda: linear_fwd(A, 128, W, 128, 1, None, H, 128, R, S, 128, 128, stats=st))
t_two, t_f = timeit(two), timeit(fused)
peak = 6543.4
print(f"B={B} R={R} T={T / 1e6:.0f} MB units={int(sl.unit_ptr[-1])}")
print(f"aggregate alone      {t_agg:7.1f} us  ({2 * T / t_agg / 1e3:.0f} GB/s)")
print(f"linear alone         {t_lin:7.1f} us  ({2 * T / t_lin / 1e3:.0f} GB/s)")
print(f"aggregate + linear   {t_two:7.1f} us  (4T: {4 * T / t_two / 1e3:.0f} GB/s)")
print(f"fused                {t_f:7.1f} us  (3T: {3 * T / t_f / 1e3:.0f} GB/s = {3 * T / t_f / 1e3 / peak:.2f} of the measured HBM peak)")

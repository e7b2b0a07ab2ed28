"""Opt-in Linear kernels (sb_set_tensor_cores(2..6): CTA pair, TMA-fed, weight resident in tensor memory; DESIGN.md
appendix) against the default tcgen05 kernel and an fp64 product, through the same C-ABI entry point.

These kernels are NOT on the default path.  Modes 2-4 were validated on a B200 through scripts/pair_check.cu
(profiles/r1z_pair_check_mode*.log); modes 5/6 have only been compiled.  The tests run when SB_EXPERIMENTAL=1 is set
(`SB_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu`) so that the round-end GPU suite exercises
exactly the code the bench runs."""
import os

import pytest
import torch

from helpers import assert_close_rel

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SB_EXPERIMENTAL") != "1", reason="opt-in kernels: set SB_EXPERIMENTAL=1")]
DEV = "cuda"


@pytest.mark.parametrize("mode", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("K,N,pro,relu,bias,transposed", [(128, 128, 2, False, True, False), (64, 96, 1, True, True, False),
                                                          (32, 32, 0, False, False, False), (128, 128, 0, False, False, True),
                                                          (96, 64, 2, True, True, True)])
def test_experimental_linear_matches_default(mode, K, N, pro, relu, bias, transposed):
    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200.functional import linear_fwd

    torch.manual_seed(K * 131 + N + mode)
    G, R = 2, 4133  # 2 x 33 tiles (ragged last tile per group; 66 tiles = 33 CTA pairs)
    x = torch.randn(G, R, K, device=DEV)
    W = (torch.randn(N, K) / K ** 0.5).to(DEV)
    Wt = W.t().contiguous()  # [K, N]: the input-gradient launches read the weight through swapped strides
    b = torch.randn(N, device=DEV) if bias else None
    pa, pc = (torch.rand(G, K) + 0.5).to(DEV), (torch.randn(G, K) * 0.3).to(DEV)
    xin = x.double()
    if pro >= 1:
        xin = xin * pa.double()[:, None, :] + pc.double()[:, None, :]
    if pro == 2:
        xin = xin.relu()
    ref = xin @ W.double().T + (b.double() if bias else 0)
    if relu:
        ref = ref.relu()

    L = _lib.lib()
    out, stats = {}, {}
    try:
        for m in (1, mode):
            L.sb_set_tensor_cores(m)
            y = torch.full((G, R, N), float("nan"), device=DEV)
            st = torch.zeros(G, 2, N, dtype=torch.float64, device=DEV)
            if transposed:
                linear_fwd(x, K, Wt, 1, N, b, y, N, R, G, K, N, pro=pro, pa=pa if pro else None, pc=pc if pro else None,
                           relu=relu, stats=st)
            else:
                linear_fwd(x, K, W, K, 1, b, y, N, R, G, K, N, pro=pro, pa=pa if pro else None, pc=pc if pro else None,
                           relu=relu, stats=st)
            torch.cuda.synchronize()
            assert L.sb_last_linear_kernel() == m, f"dispatcher fell back to kernel {L.sb_last_linear_kernel()}"
            out[m], stats[m] = y, st
    finally:
        L.sb_set_tensor_cores(1)
    assert_close_rel(out[mode].cpu(), ref.float().cpu(), 1e-5, what=f"linear mode {mode}")
    assert_close_rel(stats[mode][:, 0].cpu(), ref.sum(1).cpu(), 1e-5, floor=float(ref.abs().sum(1).max()), what="col sum")
    assert_close_rel(stats[mode][:, 1].cpu(), (ref ** 2).sum(1).cpu(), 1e-5, what="col sumsq")
    if mode in (2, 3):  # same operands, same MMA order as the default kernel
        assert torch.equal(out[mode], out[1])


@pytest.mark.parametrize("mode", [3, 4])
@pytest.mark.parametrize("pro,bias", [(0, False), (2, True)])
def test_experimental_wgrad_matches_default(mode, pro, bias):
    """TMA-fed weight gradient (csrc/wgrad_tc_tma.cu, N == K == 128) against the default tcgen05 wgrad and fp64."""
    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200.functional import linear_wgrad

    torch.manual_seed(pro * 7 + mode)
    G, R, N, K = 2, 4133, 128, 128
    gy = torch.randn(G, R, N, device=DEV)
    x = torch.randn(G, R, K, device=DEV)
    pa, pc = (torch.rand(G, K) + 0.5).to(DEV), (torch.randn(G, K) * 0.3).to(DEV)
    xin = x.double()
    if pro:
        xin = (xin * pa.double()[:, None, :] + pc.double()[:, None, :]).relu()
    ref_w = torch.einsum("grn,grk->nk", gy.double(), xin)
    ref_b = gy.double().sum((0, 1))

    L = _lib.lib()
    out = {}
    try:
        for m in (1, mode):
            L.sb_set_tensor_cores(m)
            dW = torch.full((N, K), float("nan"), device=DEV)
            db = torch.full((N,), float("nan"), device=DEV) if bias else None
            linear_wgrad(gy, N, x, K, R, G, N, K, dW, K, 1, db, pro=pro, pa=pa if pro else None, pc=pc if pro else None)
            torch.cuda.synchronize()
            assert L.sb_last_wgrad_kernel() == m, f"dispatcher fell back to kernel {L.sb_last_wgrad_kernel()}"
            out[m] = (dW, db)
    finally:
        L.sb_set_tensor_cores(1)
    assert_close_rel(out[mode][0].cpu(), ref_w.float().cpu(), 1e-5, what=f"wgrad mode {mode}")
    if bias:
        assert_close_rel(out[mode][1].cpu(), ref_b.float().cpu(), 1e-5, floor=float(gy.abs().sum((0, 1)).max()),
                         what="dbias")
    if mode == 3:  # rounded heads: same operands, same MMA order as the default kernel
        assert torch.equal(out[mode][0], out[1][0])

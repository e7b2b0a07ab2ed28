"""The two tcgen05 weight-gradient kernels behind sb_linear_wgrad - the TMA-fed wgrad_tc_tma.cu (default for N = K = 128)
and the register-fed wgrad_tc.cu (every other fast shape; forced everywhere by sb_set_tensor_cores(2)) - against each
other (same operands and deterministic reduction; the TMA-fed kernel's two MMA-issuing warps keep two partial
accumulators, so the two agree to fp32 rounding, the bias gradient bit for bit) and against an fp64 product."""
import pytest
import torch

from helpers import assert_close_rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("R", [4133, 20011])
@pytest.mark.parametrize("pro,bias,G", [(0, False, 2), (2, True, 2), (1, True, 1)])
def test_wgrad_tma_matches_register_fed(R, pro, bias, G):
    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200.functional import linear_wgrad

    torch.manual_seed(pro * 7 + R)
    N, K = 128, 128
    gy = torch.randn(G, R, N, device=DEV)
    x = torch.randn(G, R, K, device=DEV)
    pa, pc = (torch.rand(G, K) + 0.5).to(DEV), (torch.randn(G, K) * 0.3).to(DEV)
    xin = x.double()
    if pro:
        xin = xin * pa.double()[:, None, :] + pc.double()[:, None, :]
    if pro == 2:
        xin = xin.relu()
    ref_w = torch.einsum("grn,grk->nk", gy.double(), xin)
    ref_b = gy.double().sum((0, 1))

    L = _lib.lib()
    out = {}
    try:
        for mode, want in ((2, 1), (1, 3)):
            L.sb_set_tensor_cores(mode)
            dW = torch.full((N, K), float("nan"), device=DEV)
            db = torch.full((N,), float("nan"), device=DEV) if bias else None
            linear_wgrad(gy, N, x, K, R, G, N, K, dW, K, 1, db, pro=pro, pa=pa if pro else None, pc=pc if pro else None)
            torch.cuda.synchronize()
            assert L.sb_last_wgrad_kernel() == want, f"dispatcher used kernel {L.sb_last_wgrad_kernel()}, wanted {want}"
            out[mode] = (dW, db)
    finally:
        L.sb_set_tensor_cores(1)
    # sum of R*G products of O(1) terms: compare against the scale of the sum of absolute values (cancellation)
    floor = float(torch.einsum("grn,grk->nk", gy.double().abs(), xin.abs()).max()) * 1e-2
    assert_close_rel(out[1][0].cpu(), ref_w.float().cpu(), 1e-5, floor=floor, what="wgrad (TMA-fed)")
    if bias:
        assert_close_rel(out[1][1].cpu(), ref_b.float().cpu(), 1e-5, floor=float(gy.abs().sum((0, 1)).max()),
                         what="dbias")
        assert torch.equal(out[1][1], out[2][1])
    assert_close_rel(out[1][0].cpu(), out[2][0].cpu(), 2e-6, floor=floor, what="TMA-fed vs register-fed")


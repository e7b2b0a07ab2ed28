"""The fused aggregate -> Linear kernel (csrc/gin_lin_fused.cu, sb_gin_linear_fused_fwd) against the two kernels it
replaces (sb_gin_agg + sb_linear_fwd) through the C ABI - A and H bit for bit, the fp64 column statistics to rounding -
and against the CPU oracle (restate.gin_aggregate + F.linear in fp64) at 1e-5.
Reference op: MaskedGINConv.forward, Alchemy/sign_net/model_utils/masked_layers.py:74-84."""
import pytest
import torch

import restate
from helpers import assert_close_rel, rows_to_dense, slot_row_index
from signnet_basisnet_b200.synth import synth_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(B, shape, seed, S):
    from signnet_basisnet_b200.layout import GraphIndex

    d = synth_batch(B, shape, seed=seed)
    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    sl = gi.slots_all(128)
    g = torch.Generator(device=DEV).manual_seed(seed)
    X = torch.randn(S, sl.R, 128, device=DEV, generator=g)
    return d, gi, sl, X


@pytest.mark.parametrize("B,shape,S,h,eps", [(24, "zinc", 2, 128, 0.25), (200, "zinc", 2, 128, -0.1), (37, "alchemy", 1, 128, 0.0),
                                             (64, "zinc", 2, 95, 0.5), (300, "zinc", 2, 32, 0.0)])
def test_fused_matches_two_kernel_path_bit_for_bit(B, shape, S, h, eps):
    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200._lib import counted_call as call, ptr as p
    from signnet_basisnet_b200.functional import linear_fwd
    from signnet_basisnet_b200.layout import pad4
    from signnet_basisnet_b200.phi import gin_agg

    d, gi, sl, X = _setup(B, shape, 100 + B, S)
    ldh = pad4(h)
    W = (torch.randn(h, 128, generator=torch.Generator().manual_seed(h)) / 128 ** 0.5).to(DEV)
    eps_t = torch.tensor([eps], device=DEV)
    # the two-kernel path
    A0 = torch.empty_like(X)
    gin_agg(X, A0, sl, S, 128, eps=eps_t)
    H0 = torch.full((S, sl.R, ldh), float("nan"), device=DEV)
    st0 = torch.zeros(S, 2, h, dtype=torch.float64, device=DEV)
    old = _lib.lib().sb_set_small_rows(0)
    try:
        linear_fwd(A0, 128, W, 128, 1, None, H0, ldh, sl.R, S, 128, h, stats=st0)
        ref_is_tc = _lib.lib().sb_last_linear_kernel() == 1    # fewer than 4096 rows take the FFMA kernel instead
    finally:
        _lib.lib().sb_set_small_rows(old)
    # the fused kernel
    A1 = torch.full_like(X, float("nan"))
    H1 = torch.full((S, sl.R, ldh), float("nan"), device=DEV)
    st1 = torch.zeros(S, 2, h, dtype=torch.float64, device=DEV)
    assert _lib.lib().sb_set_fused_agg_linear(1) in (-1, 0, 1)
    call("sb_gin_linear_fused_fwd", p(X), p(A1), p(H1), p(st1), p(eps_t), p(W), 128, 1, 128, h, ldh, p(sl.unit_ptr),
         p(sl.unit_desc), p(gi.in_pack), p(gi.in_ptr), p(gi.in_src), sl.R, gi.B, S, 128, sl.tile_rows, 0)
    torch.cuda.synchronize()
    assert torch.equal(A1, A0), f"A differs: max {float((A1 - A0).abs().max()):.3e}"
    # same operand split as linear_tc.cu, but the fused kernel's two MMA-issuing warps accumulate K-blocks 0-1 and 2-3
    # separately (the epilogue adds the two partial sums): equal to fp32 rounding, not bit for bit
    assert_close_rel(H1.cpu(), H0.cpu(), 2e-6, what="H vs the separate Linear kernel")
    # (both kernels add 32-row partial sums in fp32 before the fp64 accumulation; the 32-row blocks differ)
    assert_close_rel(st1.cpu(), st0.cpu(), 1e-6 if ref_is_tc else 1e-5, floor=float(st0.abs().max()), what="column statistics")
    # and against the oracle's arithmetic in fp64
    idx = slot_row_index(d.batch, sl.k, True)
    for s in range(S):
        xd = rows_to_dense(X[s].cpu(), idx, 128).transpose(0, 1).double()            # [k, N, 128]
        a_ref = restate.gin_aggregate(xd, d.edge_index, torch.tensor([eps], dtype=torch.float64))
        h_ref = (a_ref @ W.double().cpu().T) * (idx >= 0).T.unsqueeze(-1)
        got = rows_to_dense(H1[s].cpu(), idx, h).transpose(0, 1)
        assert_close_rel(got, h_ref.float(), 1e-5, what=f"H vs oracle (sign {s})")


def test_unsupported_shapes_fall_back_without_error():
    from signnet_basisnet_b200 import _lib

    L = _lib.lib()
    # ld != 128 -> SB_ERR_UNSUPPORTED (3), no launch, no error message; callers run sb_gin_agg + sb_linear_fwd
    rc = L.sb_gin_linear_fused_fwd(None, None, None, None, None, None, 64, 1, 64, 64, 64, None, None, None, None, None,
                                   10, 1, 2, 64, 128, 0, None)
    assert rc == 3
    old = L.sb_set_fused_agg_linear(0)
    try:
        d, gi, sl, X = _setup(8, "zinc", 5, 2)
        rc = L.sb_gin_linear_fused_fwd(X.data_ptr(), X.data_ptr(), X.data_ptr(), None, None, X.data_ptr(), 128, 1, 128, 128,
                                       128, sl.unit_ptr.data_ptr(), sl.unit_desc.data_ptr(), gi.in_pack.data_ptr(),
                                       gi.in_ptr.data_ptr(), gi.in_src.data_ptr(), sl.R, gi.B, 2, 128, sl.tile_rows, 0, None)
        assert rc == 3   # switched off
    finally:
        L.sb_set_fused_agg_linear(1 if old != 0 else 0)


def test_phi_stack_same_result_with_and_without_fusion():
    """The whole phi stack (n_hid = 128: layers 1.. take the fused kernel) gives the same output either way."""
    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200.layout import GraphIndex, pad4
    from signnet_basisnet_b200.sign_net import GNN3d, build_phi_input

    torch.manual_seed(0)
    d = synth_batch(40, "zinc", seed=21)
    phi = GNN3d(1, 128, 3, flavour="zinc").to(DEV).train()
    gi = GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs)
    sl = gi.slots_all(pad4(128))
    x0 = build_phi_input(gi, sl, d.eigen_vectors.to(DEV))
    L = _lib.lib()
    outs = {}
    old = L.sb_set_fused_agg_linear(1)
    try:
        for f in (0, 1):
            L.sb_set_fused_agg_linear(f)
            phi2 = GNN3d(1, 128, 3, flavour="zinc").to(DEV).train()
            phi2.load_state_dict(phi.state_dict())
            xr, _ = phi2.forward_rows(x0, gi, sl.k, True)
            xr.sum().backward()
            outs[f] = (xr.detach().clone(), [p_.grad.clone() for p_ in phi2.parameters() if p_.grad is not None])
    finally:
        L.sb_set_fused_agg_linear(1 if old != 0 else 0)
    # A is bit-identical, H equal to fp32 rounding (two partial accumulators), the BatchNorm column sums are accumulated
    # in a different order: everything downstream agrees to rounding - the path's 1e-5 bar - not bit for bit
    assert_close_rel(outs[1][0], outs[0][0], 1e-5, what="phi output fused vs unfused")
    # gradients: H now differs in the last bit between the two paths, so pre-activations within rounding distance of a
    # ReLU kink fall on either side of it - the sensitivity DESIGN.md §2 measures on the fp32 oracle itself (1e-3 from its
    # own fp64 run at this depth).  The arithmetic proper is pinned at 1e-5 with the ReLU patterns imposed
    # (tests/test_gpu_signnet.py::test_phi_stack_forward_backward); this A/B check only guards against gross errors.
    for a, b in zip(outs[0][1], outs[1][1]):
        assert_close_rel(a, b, 5e-3, floor=float(b.abs().max()), what="phi gradients fused vs unfused")

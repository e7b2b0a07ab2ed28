"""Parity of the individual CUDA kernels (through the C ABI) against the CPU oracle / plain torch fp32.
Integer bookkeeping and the GIN aggregate must be bit-exact; dense contractions within 1e-5 relative."""
import numpy as np
import pytest
import torch

import restate
from helpers import assert_close_rel, dense_to_rows, rows_to_dense, slot_row_index
from signnet_basisnet_b200.synth import synth_batch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gi(d, **kw):
    from signnet_basisnet_b200.layout import GraphIndex

    return GraphIndex(d.edge_index.to(DEV), d.batch.to(DEV), d.num_graphs, **kw)


@pytest.mark.parametrize("shape,B", [("alchemy", 33), ("zinc", 17), ("zinc", 1)])
def test_graph_ptr_and_stable_csr(shape, B):
    d = synth_batch(B, shape, seed=5)
    # shuffle the edge order: the CSR must stay stable (edge-id order inside every row)
    perm = torch.randperm(d.edge_index.shape[1], generator=torch.Generator().manual_seed(0))
    d.edge_index = d.edge_index[:, perm]
    gi = _gi(d)
    gi.check()
    n = restate.graph_sizes(d.batch)
    assert np.array_equal(gi.graph_ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(n)]))
    src, dst = d.edge_index.numpy()
    N = d.batch.numel()
    for ptr, nbr, eid, key, other in ((gi.in_ptr, gi.in_src, gi.in_eid, dst, src),
                                      (gi.out_ptr, gi.out_dst, gi.out_eid, src, dst)):
        order = np.argsort(key, kind="stable")
        assert np.array_equal(ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(np.bincount(key, minlength=N))]))
        assert np.array_equal(eid.cpu().numpy()[:len(order)], order)
        assert np.array_equal(nbr.cpu().numpy()[:len(order)], other[order])


def test_bad_inputs_raise():
    from signnet_basisnet_b200.layout import GraphIndex

    d = synth_batch(4, "alchemy", seed=1)
    bad = d.batch.clone()
    bad[0], bad[-1] = bad[-1].item(), 0
    with pytest.raises(ValueError, match="sorted"):
        GraphIndex(d.edge_index.to(DEV), bad.to(DEV), 4).slots(4)
    ei = d.edge_index.clone()
    ei[0, 0] = d.batch.numel() + 3
    with pytest.raises(ValueError, match="outside"):
        GraphIndex(ei.to(DEV), d.batch.to(DEV), 4).slots(4)
    ei = d.edge_index.clone()
    ei[0, 0] = d.batch.numel() - 1  # joins graph 0 with the last graph
    with pytest.raises(ValueError, match="different graphs"):
        GraphIndex(ei.to(DEV), d.batch.to(DEV), 4).slots(4)
    with pytest.raises(ValueError, match="CUDA int64"):
        GraphIndex(d.edge_index, d.batch, 4)


@pytest.mark.parametrize("k,masked", [(5, True), (16, True), (40, True), (8, False)])
def test_slot_layout_matches_cpu(k, masked):
    d = synth_batch(21, "zinc", seed=6)
    gi = _gi(d)
    sl = gi.slots(k, masked, 128)
    idx = slot_row_index(d.batch, k, masked)
    n = restate.graph_sizes(d.batch)
    kb = np.minimum(n, k) if masked else np.full_like(n, k)
    assert sl.R == int((n * kb).sum()) == int(idx.max()) + 1
    assert np.array_equal(sl.row_ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(n * kb)]))
    assert np.array_equal(sl.vec_ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(n * n)]))
    assert sl.nmax == int(n.max())


@pytest.mark.parametrize("shape,B", [("alchemy", 19), ("zinc", 12)])
def test_dense_list_evd_bit_exact(shape, B):
    from signnet_basisnet_b200 import ops

    d = synth_batch(B, shape, seed=7)
    S_ref, V_ref = restate.dense_list_evd(d.eigen_values, d.eigen_vectors, d.batch)
    S, V, M = ops.to_dense_list_EVD(d.eigen_values.to(DEV), d.eigen_vectors.to(DEV), d.batch.to(DEV),
                                    return_mask=True)
    assert torch.equal(S.cpu(), S_ref) and torch.equal(V.cpu(), V_ref)
    assert torch.equal(M.cpu(), restate.slot_mask(d.batch, V_ref.shape[1]))


def _agg_case(d, k, masked, C, seed=0):
    idx = slot_row_index(d.batch, k, masked)
    g = torch.Generator().manual_seed(seed)
    xd = torch.randn(2, d.batch.numel(), k, C, generator=g) * (idx >= 0).unsqueeze(-1)
    return idx, xd


@pytest.mark.parametrize("C,k,masked,generic", [(128, 37, True, False), (64, 16, True, False), (96, 8, False, False),
                                                (128, 37, True, True), (1, 12, True, True), (20, 6, True, False)])
def test_gin_aggregate_bit_exact(C, k, masked, generic):
    from signnet_basisnet_b200.layout import pad4
    from signnet_basisnet_b200.phi import gin_agg

    d = synth_batch(40, "zinc", seed=8)
    # directed + shuffled edges: exercises the stable CSR and a non-symmetric neighbourhood
    keep = torch.rand(d.edge_index.shape[1], generator=torch.Generator().manual_seed(1)) < 0.8
    d.edge_index = d.edge_index[:, keep][:, torch.randperm(int(keep.sum()), generator=torch.Generator().manual_seed(2))]
    idx, xd = _agg_case(d, k, masked, C)
    eps = torch.tensor([0.37])
    ref = torch.stack([restate.gin_aggregate(xd[s].transpose(0, 1), d.edge_index, eps).transpose(0, 1) for s in (0, 1)])
    ld = pad4(C) if C > 1 else 1
    gi = _gi(d)
    sl = gi.slots(k, masked, ld)
    rows = torch.stack([dense_to_rows(xd[s], idx, ld) for s in (0, 1)]).to(DEV)
    if ld == 1:
        rows = rows.squeeze(-1).contiguous()
    out = torch.full_like(rows, float("nan"))
    gin_agg(rows, out, sl, 2, ld, eps=eps.to(DEV), force_generic=generic)
    out = out.cpu().reshape(2, -1, ld)
    for s in (0, 1):
        got = rows_to_dense(out[s], idx, C)
        assert torch.equal(got, ref[s] * (idx >= 0).unsqueeze(-1)), f"sign {s}: max diff {(got - ref[s]).abs().max()}"
    if ld > C:
        assert out[..., C:].abs().max() == 0


@pytest.mark.parametrize("C,k,masked", [(128, 37, True), (64, 16, True), (96, 8, False), (20, 6, True)])
def test_gin_aggregate_backward_mode(C, k, masked):
    """Backward use of K1: out = res + agg^T(x) with `res` aliasing `out`, plus the d-eps dot product sum(x * dotx)."""
    from signnet_basisnet_b200.layout import pad4
    from signnet_basisnet_b200.phi import gin_agg

    d = synth_batch(40, "zinc", seed=9)
    keep = torch.rand(d.edge_index.shape[1], generator=torch.Generator().manual_seed(3)) < 0.8
    d.edge_index = d.edge_index[:, keep]
    idx, xd = _agg_case(d, k, masked, C)
    _, gd = _agg_case(d, k, masked, C)
    g = torch.Generator().manual_seed(11)
    gd = torch.randn(xd.shape, generator=g) * (idx >= 0).unsqueeze(-1)
    td = torch.randn(xd.shape, generator=g) * (idx >= 0).unsqueeze(-1)
    eps = torch.tensor([-0.21])
    flipped = d.edge_index.flip(0)  # transpose of the adjacency: aggregate by source
    ref = torch.stack([gd[s] + restate.gin_aggregate(xd[s].transpose(0, 1), flipped, eps).transpose(0, 1) for s in (0, 1)])
    ref_dot = float((xd.double() * td.double()).sum())
    ld = pad4(C)
    gi = _gi(d)
    sl = gi.slots(k, masked, ld)
    to_rows = lambda t: torch.stack([dense_to_rows(t[s], idx, ld) for s in (0, 1)]).to(DEV)
    x, G, t = to_rows(xd), to_rows(gd), to_rows(td)
    for generic in (False, True):
        out = G.clone()
        dot = torch.zeros(1, dtype=torch.float64, device=DEV)
        gin_agg(x, out, sl, 2, ld, eps=eps.to(DEV), res=out, dotx=t, dot_out=dot, transpose=True, force_generic=generic)
        o = out.cpu()
        for s in (0, 1):
            got = rows_to_dense(o[s], idx, C)
            want = ref[s] * (idx >= 0).unsqueeze(-1)
            assert (got - want).abs().max() <= 1e-6 * want.abs().max(), f"generic={generic} sign {s}"
        assert abs(float(dot) - ref_dot) <= 1e-5 * max(1.0, abs(ref_dot)), (generic, float(dot), ref_dot)


@pytest.mark.parametrize("K,N,pro,relu,bias", [(128, 128, 0, False, False), (128, 128, 2, False, True),
                                               (64, 64, 2, False, True), (95, 95, 1, True, True),
                                               (1, 64, 0, False, False), (1, 1, 0, False, False),
                                               (67, 4, 2, False, True), (200, 70, 0, True, True),
                                               (6, 130, 0, False, True)])
def test_linear_fwd_and_stats(K, N, pro, relu, bias, R=333):
    """Both kernel families behind sb_linear_fwd at this size: the small-row FFMA kernel (default up to 8 192 rows) and,
    with sb_set_small_rows(0), the 128-row FFMA / tcgen05 kernels that larger problems take."""
    from signnet_basisnet_b200 import _lib

    L = _lib.lib()
    for small in (1, 0):
        old = L.sb_set_small_rows(small)
        try:
            _linear_fwd_and_stats(K, N, pro, relu, bias, R)
            if K > 1 and N > 1:
                assert (L.sb_last_linear_kernel() == 4) == bool(small), (small, L.sb_last_linear_kernel())
        finally:
            L.sb_set_small_rows(old)


def _linear_fwd_and_stats(K, N, pro, relu, bias, R):
    from signnet_basisnet_b200.functional import linear_fwd
    from signnet_basisnet_b200.layout import pad4

    torch.manual_seed(K * 131 + N)
    G = 2
    ldx, ldy = (pad4(K) if K > 1 else 1), pad4(N)
    x = torch.zeros(G, R, ldx)
    x[..., :K] = torch.randn(G, R, K)
    W = torch.randn(N, K) / K ** 0.5
    b = torch.randn(N) if bias else None
    pa, pc = torch.rand(G, K) + 0.5, torch.randn(G, K) * 0.3
    xin = x[..., :K]
    if pro >= 1:
        xin = xin * pa[:, None, :] + pc[:, None, :]
    if pro == 2:
        xin = xin.relu()
    ref = xin.double() @ W.double().T + (b.double() if bias else 0)
    if relu:
        ref = ref.relu()
    y = torch.full((G, R, ldy), float("nan"), device=DEV)
    stats = torch.zeros(G, 2, N, dtype=torch.float64, device=DEV) if K <= 128 or pro == 0 else None
    if N > 128:
        stats = None
    linear_fwd(x.to(DEV), ldx, W.to(DEV), K, 1, None if b is None else b.to(DEV), y, ldy, R, G, K, N, pro=pro,
               pa=pa.to(DEV) if pro else None, pc=pc.to(DEV) if pro else None, relu=relu, stats=stats)
    y = y.cpu()
    assert_close_rel(y[..., :N], ref.float(), 1e-5, what="linear")
    assert y[..., N:].abs().max() == 0 if ldy > N else True
    if stats is not None:
        assert_close_rel(stats[:, 0].cpu(), ref.sum(1), 1e-5, floor=float(ref.abs().sum(1).max()), what="col sum")
        assert_close_rel(stats[:, 1].cpu(), (ref ** 2).sum(1), 1e-5, what="col sumsq")


@pytest.mark.parametrize("K,N,pro", [(128, 128, 0), (128, 128, 2), (64, 64, 2), (95, 95, 1), (1, 64, 0), (64, 1, 0),
                                     (70, 200, 0), (130, 12, 0)])
def test_linear_wgrad(K, N, pro, R=777):
    from signnet_basisnet_b200.functional import linear_wgrad
    from signnet_basisnet_b200.layout import pad4

    torch.manual_seed(K * 7 + N)
    G = 2
    ldx, ldg = (pad4(K) if K > 1 else 1), (pad4(N) if N > 1 else 1)
    x = torch.zeros(G, R, ldx)
    x[..., :K] = torch.randn(G, R, K)
    gy = torch.zeros(G, R, ldg)
    gy[..., :N] = torch.randn(G, R, N)
    pa, pc = torch.rand(G, K) + 0.5, torch.randn(G, K) * 0.3
    xin = x[..., :K]
    if pro >= 1:
        xin = xin * pa[:, None, :] + pc[:, None, :]
    if pro == 2:
        xin = xin.relu()
    ref_w = torch.einsum("grn,grk->nk", gy[..., :N].double(), xin.double())
    ref_b = gy[..., :N].double().sum((0, 1))
    dW = torch.full((N, K), float("nan"), device=DEV)
    db = torch.full((N,), float("nan"), device=DEV)
    linear_wgrad(gy.to(DEV), ldg, x.to(DEV), ldx, R, G, N, K, dW, K, 1, db, pro=pro, pa=pa.to(DEV) if pro else None,
                 pc=pc.to(DEV) if pro else None)
    assert_close_rel(dW.cpu(), ref_w.float(), 1e-5, what="dW")
    assert_close_rel(db.cpu(), ref_b.float(), 1e-5, what="db")


@pytest.mark.parametrize("K,N,pro,relu,bias", [(1, 128, 0, False, False), (1, 64, 1, True, True), (1, 95, 2, False, True),
                                               (128, 1, 0, False, False), (95, 1, 2, True, True),
                                               (128, 128, 2, False, True), (64, 128, 0, True, False),
                                               (95, 95, 1, True, True), (96, 70, 2, False, True), (20, 128, 0, False, False)])
def test_linear_fwd_streaming_sizes(K, N, pro, relu, bias):
    """Row counts above the small-problem threshold: the rank-1 / row-dot kernels of the first phi layer
    (csrc/linear_rank1.cu) and the tcgen05 path (csrc/linear_tc.cu), including a ragged last tile."""
    test_linear_fwd_and_stats(K, N, pro, relu, bias, R=2999)


@pytest.mark.parametrize("K,N,pro", [(1, 128, 0), (1, 64, 2), (1, 95, 1), (128, 128, 2), (95, 95, 1), (64, 128, 0),
                                     (70, 96, 2)])
def test_linear_wgrad_streaming_sizes(K, N, pro):
    test_linear_wgrad(K, N, pro, R=2999)


@pytest.mark.parametrize("M,C,relu,res,training", [(2922, 64, True, True, True), (333, 95, True, False, True),
                                                   (8192, 128, False, False, True), (1, 6, True, True, True),
                                                   (700, 67, True, True, False)])
def test_bn_act_small_kernel_matches_streaming_and_torch(M, C, relu, res, training):
    """sb_bn_act_fwd / sb_bn_act_bwd: the one-kernel path (<= 8 192 rows) is bit-identical to the three streaming kernels
    it replaces, and both are nn.BatchNorm1d (+ ReLU + residual) within 1e-5 (forward, input / affine gradients,
    running buffers)."""
    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200.functional import batch_norm_act
    from signnet_basisnet_b200.layout import pad4

    torch.manual_seed(M + C)
    ld = pad4(C)
    x0 = torch.zeros(M, ld)
    x0[:, :C] = torch.randn(M, C) * 1.7 + 0.3
    r0 = torch.zeros(M, ld)
    r0[:, :C] = torch.randn(M, C)
    g0 = torch.zeros(M, ld)
    g0[:, :C] = torch.randn(M, C)
    bn_ref = torch.nn.BatchNorm1d(C)
    with torch.no_grad():
        bn_ref.weight.copy_(torch.rand(C) + 0.5)
        bn_ref.bias.copy_(torch.randn(C) * 0.2)
        bn_ref.running_mean.copy_(torch.randn(C) * 0.1)
        bn_ref.running_var.copy_(torch.rand(C) + 0.5)
    bn_ref.train(training)
    sd = {k: v.clone() for k, v in bn_ref.state_dict().items()}

    def run(small):
        old = _lib.lib().sb_set_small_bn(small)
        try:
            bn = torch.nn.BatchNorm1d(C).to(DEV)
            bn.load_state_dict(sd)
            bn.train(training)
            x = x0.to(DEV).requires_grad_(True)
            r = r0.to(DEV).requires_grad_(True) if res else None
            out = batch_norm_act(x, bn, training, relu=relu, res=r)
            out.backward(g0.to(DEV))
            torch.cuda.synchronize()
            return [out.detach().cpu(), x.grad.cpu(), bn.weight.grad.cpu(), bn.bias.grad.cpu(), bn.running_mean.cpu(),
                    bn.running_var.cpu()] + ([r.grad.cpu()] if res else [])
        finally:
            _lib.lib().sb_set_small_bn(old)

    a, b = run(1), run(0)
    for i, (u, v) in enumerate(zip(a, b)):
        assert torch.equal(u, v), (i, float((u - v).abs().max()))
    if M > 1:
        xr = x0[:, :C].double().requires_grad_(True)
        bn64 = torch.nn.BatchNorm1d(C).double()
        bn64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
        bn64.train(training)
        y = bn64(xr)
        y = y.relu() if relu else y
        if res:
            y = y + r0[:, :C].double()
        y.backward(g0[:, :C].double())
        assert_close_rel(a[0][:, :C].double(), y.detach(), 1e-5, what="bn_act out")
        assert_close_rel(a[1][:, :C].double(), xr.grad, 1e-5, what="bn_act dx")
        assert_close_rel(a[2].double(), bn64.weight.grad, 1e-5, what="bn_act dgamma")
        assert_close_rel(a[3].double(), bn64.bias.grad, 1e-5, what="bn_act dbeta")
        assert_close_rel(a[4].double(), bn64.running_mean, 1e-5, what="running_mean")
        assert_close_rel(a[5].double(), bn64.running_var, 1e-5, what="running_var")
    assert float(a[0][:, C:].abs().max()) == 0.0 if ld > C else True


@pytest.mark.parametrize("R,K,N", [(20011, 128, 128), (9000, 64, 96), (9000, 95, 95)])
def test_linear_accumulate_on_the_tensor_core_path(R, K, N):
    """y += x W^T (the input-gradient accumulation of rho's fused blocks): the FAST tcgen05 template adds the old y in its
    write-back stage (K, N multiples of 32), the generic template in registers; both against fp64."""
    from signnet_basisnet_b200 import _lib
    from signnet_basisnet_b200.functional import linear_fwd
    from signnet_basisnet_b200.layout import pad4

    torch.manual_seed(R + K)
    ldx, ldy = pad4(K), pad4(N)
    x = torch.zeros(R, ldx, device=DEV)
    x[:, :K] = torch.randn(R, K, device=DEV)
    W = (torch.randn(N, K) / K ** 0.5).to(DEV)
    y0 = torch.zeros(R, ldy, device=DEV)
    y0[:, :N] = torch.randn(R, N, device=DEV)
    ref = y0[:, :N].double() + x[:, :K].double() @ W.double().T
    y = y0.clone()
    linear_fwd(x, ldx, W, K, 1, None, y, ldy, R, 1, K, N, accumulate=True)
    torch.cuda.synchronize()
    assert _lib.lib().sb_last_linear_kernel() == 1
    assert_close_rel(y[:, :N].double().cpu(), ref.cpu(), 1e-5, what="y += x W^T")
    if ldy > N:
        assert float(y[:, N:].abs().max()) == 0.0

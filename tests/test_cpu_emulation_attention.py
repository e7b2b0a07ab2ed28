"""Serial CPU emulation of csrc/graph_attention.cu (edge-modulated sparse attention of the graph-Transformer predictor)
against the fp64 oracle restate.sparse_attention, which is pinned against the reference's TransformerNet
(tests/test_oracle_vs_reference.py).  See tests/cpu_emulation.py for why a thread-by-thread run is faithful."""
import ctypes

import pytest
import torch

import cpu_emulation
import restate
from cpu_emulation import stable_csr
from signnet_basisnet_b200.synth import synth_batch

WRAPPERS = r"""
extern "C" void emu_fwd(const float* Q, const float* K, const float* Ef, const float* V, const int32_t* in_ptr,
                        const int32_t* in_src, const int32_t* in_eid, long long N, int H, int d, long long ld, float* out,
                        float* araw, float* z) {
  LAUNCH(edge_attention_fwd_kernel, (N * H + 255) / 256, Q, K, Ef, V, in_ptr, in_src, in_eid, N, H, d, ld, out, araw, z)
}
extern "C" void emu_bwd(const float* dout, const float* out, const float* Q, const float* K, const float* Ef, const float* V,
                        const float* araw, const float* z, const int32_t* in_ptr, const int32_t* in_src,
                        const int32_t* in_eid, const int32_t* out_ptr, const int32_t* out_eid, long long N, int H, int d,
                        long long ld, float* dQ, float* dK, float* dE, float* dV, float* dKe, float* dVe) {
  LAUNCH(edge_attention_bwd_dst_kernel, (N * H + 255) / 256, dout, out, Q, K, Ef, V, araw, z, in_ptr, in_src, in_eid, N, H, d,
         ld, dQ, dE, dKe, dVe)
  LAUNCH(edge_attention_bwd_src_kernel, (N * ld + 255) / 256, dKe, dVe, out_ptr, out_eid, N, ld, dK, dV)
}
"""


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    lib, n = cpu_emulation.build(str(tmp_path_factory.mktemp("emu_att")), "graph_attention.cu",
                                 [r"__global__ void __launch_bounds__\(256\) edge_attention_\w+"],
                                 WRAPPERS, defines="#define GA_MAXD 32")
    assert n == 3
    return lib


@pytest.mark.parametrize("B,H,d,ld,scale", [(7, 4, 4, 16, 1.0), (5, 8, 8, 64, 1.0), (4, 2, 5, 12, 3.0)])
def test_edge_attention_source_emulated(emu, B, H, d, ld, scale):
    """scale = 3 drives part of the scores outside [-5, 5] so that the clamp (and its zero gradient) is exercised."""
    g_ = synth_batch(B, "zinc", seed=50 + B)
    N, E, C = g_.batch.numel(), g_.edge_index.shape[1], H * d
    src, dst = g_.edge_index
    in_ptr, in_src, in_eid = stable_csr(dst, src, N)
    out_ptr, _, out_eid = stable_csr(src, dst, N)
    gen = torch.Generator().manual_seed(4)

    def pad(t):
        o = torch.zeros(t.shape[0], ld)
        o[:, :C] = t
        return o.contiguous()

    Qr, Kr, Vr = (torch.randn(N, C, generator=gen) * scale for _ in range(3))
    Er = torch.randn(E, C, generator=gen) * scale
    Q, K, V, Ef = pad(Qr), pad(Kr), pad(Vr), pad(Er)
    nan = float("nan")
    out, araw, z = torch.full((N, ld), nan), torch.full((E, H), nan), torch.full((N, H), nan)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    LL, I = ctypes.c_longlong, ctypes.c_int
    emu.emu_fwd(P(Q), P(K), P(Ef), P(V), P(in_ptr), P(in_src), P(in_eid), LL(N), I(H), I(d), LL(ld), P(out), P(araw), P(z))
    assert not torch.isnan(out).any() and not torch.isnan(araw).any() and not torch.isnan(z).any()

    w = torch.randn(N, C, generator=gen)

    def oracle(dt):
        q, k, e, v = (t.to(dt).clone().requires_grad_(True) for t in (Qr, Kr, Er, Vr))
        o = restate.sparse_attention(q, k, e, v, src, dst, H)
        (o * w.to(dt)).sum().backward()
        return [t.double() for t in (o.detach(), q.grad, k.grad, e.grad, v.grad)]

    o32, o64 = oracle(torch.float32), oracle(torch.float64)
    tol = [max(2e-5, 3.0 * float((a - b).abs().max())) for a, b in zip(o32, o64)]   # yardstick: the same formula in fp32
    if scale > 1:
        a64 = ((Kr.double().reshape(N, H, d)[src] * Qr.double().reshape(N, H, d)[dst]) / d ** 0.5 * Er.double().reshape(E, H, d)).sum(-1)
        assert bool((a64.abs() > 5).any()) and bool((a64.abs() < 5).any())
    assert float((out[:, :C].double() - o64[0]).abs().max()) <= tol[0]
    assert float(out[:, C:].abs().sum()) == 0

    dout = pad(w)
    dQ, dK, dV = (torch.full((N, ld), nan) for _ in range(3))
    dE, dKe, dVe = (torch.full((E, ld), nan) for _ in range(3))
    emu.emu_bwd(P(dout), P(out), P(Q), P(K), P(Ef), P(V), P(araw), P(z), P(in_ptr), P(in_src), P(in_eid), P(out_ptr), P(out_eid),
                LL(N), I(H), I(d), LL(ld), P(dQ), P(dK), P(dE), P(dV), P(dKe), P(dVe))
    for name, got, want, t in (("dQ", dQ, o64[1], tol[1]), ("dK", dK, o64[2], tol[2]), ("dE", dE, o64[3], tol[3]),
                               ("dV", dV, o64[4], tol[4])):
        assert not torch.isnan(got).any(), name
        err = float((got[:, :C].double() - want).abs().max())
        assert err <= t, (name, err, t)
        assert float(got[:, C:].abs().sum()) == 0, name

"""Serial CPU emulation of csrc/gated.cu (the edge-gated aggregate of the GatedGCN predictor) against an fp64 statement
of GraphPrediction/layers/gatedgcn_layer.py:48-54.

The three kernels have no inter-thread communication (no shuffles, barriers, shared memory or atomics), so running
their SOURCE TEXT thread by thread under a tiny prelude that defines threadIdx / __ldg / float4 ... is a faithful
execution of their indexing and arithmetic.  This is test infrastructure for a kernel file that was written after the
round's GPU budget was spent; it is not a CPU path of the product (nothing in signnet_basisnet_b200/ can reach it)."""
import ctypes

import pytest
import torch

import cpu_emulation
from cpu_emulation import stable_csr as _stable_csr
from signnet_basisnet_b200.synth import synth_batch

WRAPPERS = r"""
extern "C" void emu_fwd(const float* Ah, const float* Bh, const float* Dh, const float* Eh, const float* Ce,
                        const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid, long long N, int ld,
                        float* e_out, float* h_out, float* ss, float* ssh) {
  LAUNCH(gated_agg_fwd_kernel, (N * 32 + 255) / 256, Ah, Bh, Dh, Eh, Ce, in_ptr, in_src, in_eid, N, ld, e_out, h_out, ss, ssh)
}
extern "C" void emu_bwd(const float* dh, const float* de, const float* Bh, const float* e_new, const float* ss,
                        const float* ssh, const int64_t* ei, const int32_t* in_ptr, const int32_t* in_eid,
                        const int32_t* out_ptr, const int32_t* out_dst, const int32_t* out_eid, long long N, long long E,
                        int ld, float* dBh, float* dDh, float* dEh, float* dCe) {
  LAUNCH(gated_agg_bwd_edge_kernel, (E * 32 + 255) / 256, dh, de, Bh, e_new, ss, ssh, ei, ei + E, E, ld, dCe)
  LAUNCH(gated_agg_bwd_node_kernel, (N * 32 + 255) / 256, dh, e_new, ss, dCe, in_ptr, in_eid, out_ptr, out_dst, out_eid, N,
         ld, dBh, dDh, dEh)
}
extern "C" void emu_canonical(const float* pe, long long ldp, const int32_t* gp, long long B, int k, float* out,
                              long long ldo) {
  LAUNCH(canonical_sign_kernel, (B * k * 32 + 255) / 256, pe, ldp, gp, B, k, out, ldo)
}
"""


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    lib, n = cpu_emulation.build(str(tmp_path_factory.mktemp("emu")), "gated.cu",
                                 [r"__device__ __forceinline__ float gt_sigmoid",
                                  r"__global__ void __launch_bounds__\(256\) (?:gated_agg|canonical_sign)_\w+"], WRAPPERS)
    assert n == 5
    return lib


@pytest.mark.parametrize("B,C,ld,with_de", [(9, 18, 20, True), (4, 128, 128, True), (6, 67, 68, False)])
def test_gated_aggregate_source_emulated(emu, B, C, ld, with_de):
    d = synth_batch(B, "zinc", seed=31 + B)
    N, E = d.batch.numel(), d.edge_index.shape[1]
    src, dst = d.edge_index
    in_ptr, in_src, in_eid = _stable_csr(dst, src, N)
    out_ptr, out_dst, out_eid = _stable_csr(src, dst, N)
    gen = torch.Generator().manual_seed(1)

    def pad(t):
        o = torch.zeros(t.shape[0], ld)
        o[:, :C] = t
        return o.contiguous()

    ins = [torch.randn(N, C, generator=gen) for _ in range(4)] + [torch.randn(E, C, generator=gen)]
    Ah, Bh, Dh, Eh, Ce = (pad(t) for t in ins)
    nan = float("nan")
    e_out, h_out = torch.full((E, ld), nan), torch.full((N, ld), nan)
    ss, ssh = torch.full((N, ld), nan), torch.full((N, ld), nan)
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    LL, I = ctypes.c_longlong, ctypes.c_int
    emu.emu_fwd(P(Ah), P(Bh), P(Dh), P(Eh), P(Ce), P(in_ptr), P(in_src), P(in_eid), LL(N), I(ld), P(e_out), P(h_out), P(ss),
                P(ssh))
    ref_in = [t.double().clone().requires_grad_(True) for t in ins]
    A, Bm, D, Em, Cm = ref_in
    e_ref = (D[src] + Em[dst]) + Cm
    sg = torch.sigmoid(e_ref)
    z = lambda: torch.zeros(N, C, dtype=torch.float64)
    h_ref = A + z().index_add(0, dst, Bm[src] * sg) / (z().index_add(0, dst, sg) + 1e-6)
    assert not torch.isnan(h_out).any() and not torch.isnan(e_out).any()
    torch.testing.assert_close(h_out[:, :C].double(), h_ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(e_out[:, :C].double(), e_ref.detach(), rtol=1e-5, atol=1e-5)
    assert float(h_out[:, C:].abs().sum()) == 0 and float(e_out[:, C:].abs().sum()) == 0   # padding columns stay 0

    wh, we = torch.randn(N, C, generator=gen), torch.randn(E, C, generator=gen)
    loss = (h_ref * wh.double()).sum() + ((e_ref * we.double()).sum() if with_de else 0)
    loss.backward()
    dh, de = pad(wh), (pad(we) if with_de else None)
    dB, dD, dE = (torch.full((N, ld), nan) for _ in range(3))
    dC = torch.full((E, ld), nan)
    ei = d.edge_index.contiguous()
    emu.emu_bwd(P(dh), P(de), P(Bh), P(e_out), P(ss), P(ssh), P(ei), P(in_ptr), P(in_eid), P(out_ptr), P(out_dst), P(out_eid),
                LL(N), LL(E), I(ld), P(dB), P(dD), P(dE), P(dC))
    for name, got, want in (("dBh", dB, Bm.grad), ("dDh", dD, D.grad), ("dEh", dE, Em.grad), ("dCe", dC, Cm.grad)):
        assert not torch.isnan(got).any(), name
        torch.testing.assert_close(got[:, :C].double(), want, rtol=2e-5, atol=2e-5, msg=name)
        assert float(got[:, C:].abs().sum()) == 0, name


def test_canonical_sign_source_emulated(emu):
    """canonical_sign_kernel (one warp per (graph, column), lane-0 decision broadcast) against oracle/restate.handle_lap,
    which is bit-exact against train_ZINC_graph_regression.py:26-42."""
    import restate

    d = synth_batch(13, "zinc", seed=8, k_dgl=8)
    pe = d.pos_enc.contiguous()
    n = torch.as_tensor(d.num_nodes_per_graph)
    gp = torch.zeros(n.numel() + 1, dtype=torch.int32)
    gp[1:] = torch.cumsum(n, 0).to(torch.int32)
    out = torch.full_like(pe, float("nan"))
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    emu.emu_canonical(P(pe), ctypes.c_longlong(pe.stride(0)), P(gp), ctypes.c_longlong(n.numel()), ctypes.c_int(pe.shape[1]),
                      P(out), ctypes.c_longlong(out.stride(0)))
    ref = restate.handle_lap(pe.clone(), d.num_nodes_per_graph, "canonical")
    assert torch.equal(out, ref)
    assert bool((ref != pe).any())   # the case really flips some columns

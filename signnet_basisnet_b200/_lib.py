"""ctypes binding of libsignnet_b200.so (the C ABI declared in include/signnet_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the product path raises.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsignnet_b200.so")

# p = device pointer (or NULL), l = int64, i = int32, f = float
_SIGNATURES = {
    "sb_graph_ptr": "plipp" + "p",
    "sb_slot_layout": "piiii" + "pppp" + "p",
    "sb_agg_units": "piiii" + "p" + "p",
    "sb_build_csr": "pllp" + "pppppp" + "plp" + "p",
    "sb_phi_input_ragged": "ppppp" + "liil" + "p" + "p",
    "sb_phi_input_dense": "plppp" + "liil" + "p" + "p",
    "sb_slot_eigval": "pppp" + "lii" + "p" + "p",
    "sb_dense_list_evd": "ppppp" + "li" + "ppp" + "p",
    "sb_rows_to_dense": "pll" + "i" + "ppp" + "l" + "iii" + "p" + "p",
    "sb_dense_to_rows": "pl" + "ii" + "ppp" + "l" + "iii" + "l" + "p" + "p",
    "sb_agg_unit_desc": "ppppp" + "iiii" + "pl" + "p",
    "sb_pack_neighbours": "pppp" + "l" + "p" + "p",
    "sb_gin_agg": "ppppp" + "p" + "ppppppp" + "l" + "iiiiiii" + "p",
    "sb_gin_linear_fused_fwd": "pppppp" + "ll" + "ii" + "l" + "ppppp" + "l" + "iiiii" + "p",
    "sb_gine_stack_fwd": "pipp" + "i" + "ff" + "p",
    "sb_gine_stack_bwd": "pippp" + "i" + "p",
    "sb_phi_stack_fwd": "ppipp" + "ii" + "ff" + "p",
    "sb_phi_stack_bwd": "ppippp" + "ii" + "p",
    "sb_linear_fwd": "pl" + "pll" + "p" + "pl" + "l" + "iii" + "ipp" + "i" + "p" + "i" + "p",
    "sb_linear_wgrad": "pl" + "pl" + "l" + "iii" + "ipp" + "pll" + "p" + "i" + "p" + "p",
    "sb_col_stats": "pll" + "ii" + "p" + "p",
    "sb_bn_finalize": "pl" + "ii" + "pppp" + "ff" + "i" + "ppp" + "p",
    "sb_affine_act_res": "ppppp" + "ll" + "iii" + "p",
    "sb_bn_bwd_reduce": "pppppp" + "ll" + "iii" + "p" + "p",
    "sb_bn_bwd_finalize": "pl" + "ii" + "pp" + "ii" + "ppp" + "p",
    "sb_affine2": "ppppppp" + "ll" + "ii" + "p",
    "sb_gine_agg_fwd": "pppppp" + "li" + "p" + "p",
    "sb_gine_agg_bwd": "pppp" + "pppp" + "lli" + "ppp" + "p",
    "sb_gated_agg_fwd": "ppppp" + "ppp" + "li" + "pppp" + "p",
    "sb_gated_agg_bwd": "pppppp" + "p" + "ppppp" + "lli" + "pppp" + "p",
    "sb_pna_agg_fwd": "pppp" + "ppp" + "lii" + "lll" + "f" + "p" + "p",
    "sb_pna_agg_bwd": "pppp" + "ppp" + "pp" + "lii" + "lll" + "f" + "pppp" + "p",
    "sb_row_scale": "pp" + "ll" + "p" + "p",
    "sb_leaky_relu": "pp" + "l" + "f" + "p" + "p",
    "sb_edge_attention_fwd": "pppp" + "ppp" + "lii" + "l" + "ppp" + "p",
    "sb_edge_attention_bwd": "pppppp" + "pp" + "ppp" + "pp" + "lii" + "l" + "pppppp" + "p",
    "sb_canonical_sign": "plp" + "li" + "pl" + "p",
    "sb_segment_pool_fwd": "plp" + "iii" + "pl" + "p",
    "sb_segment_pool_bwd": "plpp" + "lii" + "pl" + "p",
    "sb_embedding_fwd": "plp" + "iil" + "pl" + "ip" + "p",
    "sb_embedding_bwd": "plpl" + "iil" + "pp" + "p",
    "sb_attention_fwd": "pppl" + "ppp" + "l" + "iiiii" + "ff" + "l" + "p" + "p",
    "sb_attention_bwd": "ppppl" + "ppp" + "l" + "iiiii" + "ff" + "l" + "ppp" + "p",
    "sb_layernorm_fwd": "pppp" + "ll" + "i" + "f" + "ppp" + "p",
    "sb_layernorm_bwd": "pppp" + "ll" + "i" + "pp" + "p",
    "sb_relu_bwd": "pppl" + "p",
    "sb_bn_apply_fwd": "ppl" + "ii" + "pppp" + "ff" + "ii" + "pp" + "ll" + "ppp" + "p",
    "sb_bn_apply_bwd": "ppppppp" + "li" + "ppp" + "ll" + "ii" + "p",
    "sb_bn_act_fwd": "pll" + "ii" + "pppp" + "ff" + "ii" + "pp" + "pppp" + "p",
    "sb_bn_act_bwd": "pppppp" + "ll" + "iiii" + "ppp" + "pp" + "p",
    "sb_laplacian_evd": "pppp" + "ii" + "ppp" + "p",
    "sb_ign2to1_ops_factors": "plipiipi" + "p",
    "sb_ign2to1_ops_projectors": "piipip" + "p",
    "sb_slot_sum_fwd": "pll" + "i" + "ppp" + "l" + "iii" + "pl" + "i" + "p",
    "sb_slot_sum_bwd": "pl" + "pll" + "i" + "ppp" + "l" + "iii" + "i" + "p",
}
_CT = {"p": ctypes.c_void_p, "l": ctypes.c_int64, "i": ctypes.c_int32, "f": ctypes.c_float}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} not found: build it with `python -m signnet_basisnet_b200.build` "
                "(there is no CPU or PyTorch fallback for the SignNet hot path)")
        L = ctypes.CDLL(LIB_PATH)
        L.sb_last_error.restype = ctypes.c_char_p
        L.sb_abi_version.restype = ctypes.c_int
        L.sb_device_sm_count.restype = ctypes.c_int
        L.sb_gin_agg_tile_rows.restype = ctypes.c_int
        L.sb_gin_agg_tile_rows.argtypes = [ctypes.c_int32]
        L.sb_linear_wgrad_workspace_floats.restype = ctypes.c_int64
        L.sb_set_tensor_cores.restype = ctypes.c_int
        L.sb_set_tensor_cores.argtypes = [ctypes.c_int32]
        L.sb_set_fused_agg_linear.restype = ctypes.c_int
        L.sb_set_fused_agg_linear.argtypes = [ctypes.c_int32]
        L.sb_set_small_rows.restype = ctypes.c_int
        L.sb_set_small_rows.argtypes = [ctypes.c_int32]
        L.sb_set_attention_mma.restype = ctypes.c_int
        L.sb_set_attention_mma.argtypes = [ctypes.c_int32]
        L.sb_set_small_bn.restype = ctypes.c_int
        L.sb_set_small_bn.argtypes = [ctypes.c_int32]
        L.sb_last_linear_kernel.restype = ctypes.c_int
        L.sb_last_wgrad_kernel.restype = ctypes.c_int
        L.sb_embedding_bwd_workspace_floats.restype = ctypes.c_int64
        L.sb_embedding_bwd_workspace_floats.argtypes = [ctypes.c_int32, ctypes.c_int32]
        for name, sig in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [_CT[c] for c in sig]
        _lib = L
    return _lib


def exported_symbols():
    return sorted(list(_SIGNATURES) + ["sb_last_error", "sb_abi_version", "sb_device_sm_count", "sb_gin_agg_tile_rows",
                                       "sb_linear_wgrad_workspace_floats", "sb_embedding_bwd_workspace_floats",
                                       "sb_set_tensor_cores", "sb_set_small_rows", "sb_set_small_bn", "sb_set_attention_mma",
                                       "sb_set_fused_agg_linear", "sb_last_linear_kernel",
                                       "sb_last_wgrad_kernel"])


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_raw_stream = torch._C._cuda_getCurrentRawStream   # (device index) -> cudaStream_t of torch's current stream


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device.  (torch.cuda.current_stream().cuda_stream costs
    ~15 us of Python per call - 5 ms per step at ~340 C-ABI calls, scripts/host_probe.py - the raw accessor < 1 us.)"""
    return _raw_stream(torch._C._cuda_getDevice())


_fn_cache: dict = {}


def call(name, *args):
    """Invoke an entry point on torch's current stream; raises RuntimeError(sb_last_error()) on failure."""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(lib(), name)
    rc = fn(*args, _raw_stream(torch._C._cuda_getDevice()))
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib().sb_last_error().decode()}")


def try_call(name, *args):
    """Like counted_call for entry points that may answer SB_ERR_UNSUPPORTED (3) without an error: returns True if the
    call ran, False if the caller must take its fallback; any other non-zero code raises."""
    global launch_count
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(lib(), name)
    if _profile is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = fn(*args, _raw_stream(torch._C._cuda_getDevice()))
    if rc == 3:
        return False
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib().sb_last_error().decode()}")
    launch_count += 1
    if _profile is not None:
        e1.record()
        _profile.append((_profile_tag(name, args), e0, e1))
    return True


launch_count = 0  # number of C-ABI calls issued (each enqueues >= 1 kernel); bench.py reports it
_profile = None   # None = off; else list of (tag, start_event, end_event) on torch's current stream


def counted_call(name, *args):
    global launch_count
    launch_count += 1
    if _profile is None:
        call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, *args)
    e1.record()
    _profile.append((_profile_tag(name, args), e0, e1))


def _profile_tag(name, args):
    if name == "sb_gin_agg":  # (.. 13 pointers, R, B, k, masked, S, ld, tile_rows, force_generic): split by row width / path
        return f"sb_gin_agg[ld={args[18]}{',generic' if args[20] else ''}{',bwd' if args[2] or args[3] else ''}]"
    if name == "sb_linear_fwd":    # (x, ldx, w, rs, cs, bias, y, ldy, R, G, K, N, pro, ...)
        return f"sb_linear_fwd[K={args[10]},N={args[11]},rows={args[8] * args[9]}]"
    if name == "sb_linear_wgrad":  # (gy, ldg, x, ldx, R, G, N, K, ...)
        return f"sb_linear_wgrad[N={args[6]},K={args[7]},rows={args[4] * args[5]}]"
    return name


def profile_start():
    global _profile
    _profile = []


def profile_stop():
    """-> {tag: (calls, total_ms)} measured with CUDA events on the launching stream."""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    out = {}
    for tag, e0, e1 in rec or []:
        c, t = out.get(tag, (0, 0.0))
        out[tag] = (c + 1, t + e0.elapsed_time(e1))
    return out

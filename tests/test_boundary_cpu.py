"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/signnet_b200.h
declares with the argument lists the ctypes binding uses; module constructors/state_dict keys match the reference;
the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import ref_loader
from signnet_basisnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    hdr = open(os.path.join(ROOT, "include", "signnet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return re.findall(r"\b(?:int|int64_t|const char\*)\s+(sb_\w+)\(([^;]*?)\);", hdr, re.S)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with `python -m signnet_basisnet_b200.build`"
    L = ctypes.CDLL(_lib.LIB_PATH)
    decls = _header_decls()
    assert len(decls) >= 35
    for name, _ in decls:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert set(n for n, _ in decls) == set(_lib.exported_symbols())
    assert L.sb_abi_version() == 1


def test_ctypes_signatures_match_header():
    for name, args in _header_decls():
        sig = _lib._SIGNATURES.get(name)
        if sig is None:
            continue
        kinds = "".join("p" if "*" in a else "l" if "int64_t" in a else "i" if "int32_t" in a else
                        "f" if "float" in a else "?" for a in args.split(","))
        assert kinds == sig, f"{name}: header {kinds} vs binding {sig}"


def test_no_cpu_fallback():
    from signnet_basisnet_b200.layout import GraphIndex
    from signnet_basisnet_b200.sign_net import SignNet
    from signnet_basisnet_b200.synth import synth_batch

    d = synth_batch(2, "alchemy", seed=0)
    with pytest.raises(ValueError, match="CUDA"):
        GraphIndex(d.edge_index, d.batch, 2)
    with pytest.raises(ValueError, match="CUDA"):
        SignNet(8, 1, nl_rho=0)(d)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "signnet_basisnet_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(import|from)\s+(restate|ref_loader|oracle)\b", src, re.M), fn
            assert "sys.path" not in src, fn


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("flavour", ["alchemy", "zinc"])
def test_state_dict_keys_and_shapes_match_reference(flavour):
    import importlib
    import sys

    from signnet_basisnet_b200.sign_net import SignNetGNN

    if flavour == "alchemy":
        ref = ref_loader.alchemy().SignNetGNN(6, 4, n_hid=16, n_out=12, nl_signnet=3, nl_gnn=2)
        mine = SignNetGNN(6, 4, n_hid=16, n_out=12, nl_signnet=3, nl_gnn=2)
    else:
        sys.path.insert(0, os.path.join(ref_loader.REF_ROOT, "GINESignNetPyG"))
        ref_loader.alchemy()  # puts the shim on sys.path
        ref = importlib.import_module("core.sign_net").SignNetGNN(None, None, 16, 1, 3, 2)
        mine = SignNetGNN(None, None, 16, 1, 3, 2, flavour="zinc")
    a, b = mine.state_dict(), ref.state_dict()
    assert set(a) == set(b), (sorted(set(a) - set(b))[:5], sorted(set(b) - set(a))[:5])
    for k in a:
        assert a[k].shape == b[k].shape, k
    mine.load_state_dict(b)  # a reference checkpoint loads unchanged
    assert sum(p.numel() for p in mine.parameters()) == sum(p.numel() for p in ref.parameters())


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_alchemy_config_parameter_count():
    """SURVEY §8c [probe]: SignNetGNN(6,4,n_hid=64,n_out=12,nl_signnet=8,nl_gnn=16) has 326 527 parameters."""
    from signnet_basisnet_b200.sign_net import SignNetGNN

    assert sum(p.numel() for p in SignNetGNN(6, 4, 64, 12, 8, 16).parameters()) == 326527

"""CPU: the C-ABI library loads and exports every symbol include/avlmaps_b200.h declares; compute
calls fail loudly without a device (there is no CPU fallback to fall into)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from avlmaps_b200 import _lib as L

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "avlmaps_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(avl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(L.EXPORTS) == names


def test_version_and_error_string():
    lib = L.load()
    assert lib.avl_version() == 100
    assert isinstance(lib.avl_last_error(), bytes)


def test_argument_errors_need_no_device():
    lib = L.load()
    n = C.c_int(-1)
    assert lib.avl_device_count(C.byref(n)) == 0 and n.value >= 0
    out = C.c_void_p()
    feat = np.zeros((4, 8), np.float32)
    # invalid arguments are rejected before any CUDA call
    assert lib.avl_map_create(L.np_ptr(feat), -1, 8, 0, None, C.byref(out)) == 2
    assert lib.avl_map_create(L.np_ptr(feat), 4, 0, 0, None, C.byref(out)) == 2
    assert b"dim" in lib.avl_last_error()
    assert lib.avl_builder_create(None, C.byref(out)) == 2
    # the peer-memory exchange: shape checks come before any CUDA call
    assert lib.avl_p2p_create(0, 0, 256, 16, C.byref(out)) == 2 and lib.avl_p2p_create(3, 2, 256, 16, C.byref(out)) == 2
    assert lib.avl_p2p_create(0, 16, 256, 128, C.byref(out)) == 2 and b"1024" in lib.avl_last_error()
    assert lib.avl_p2p_handle_bytes() == 64 and lib.avl_p2p_destroy(None) == 0


def test_no_cpu_fallback_without_device():
    if L.device_count() > 0:
        pytest.skip("a GPU is present")
    from avlmaps_b200.engine import DeviceBuilder, DeviceMap

    with pytest.raises(L.AvlError):
        DeviceMap(np.zeros((4, 8), np.float32))
    with pytest.raises(L.AvlError):
        DeviceBuilder(8, 4, 0.05, 8)
    out = C.c_void_p()
    assert L.load().avl_p2p_create(0, 1, 256, 16, C.byref(out)) == 1 and not out.value      # AVL_ERR_CUDA, no handle


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under avlmaps_b200/ may import, include, dlopen or exec it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|importlib.*oracle|liboracle|oracle/_lib|oracle\.avl_oracle", re.M)
    for p in (ROOT / "avlmaps_b200").rglob("*.py"):
        assert not pat.search(p.read_text()), f"{p} uses oracle/"
    for p in (ROOT / "avlmaps_b200" / "csrc").glob("*"):
        assert not re.search(r"#include\s*[<\"].*oracle", p.read_text()), p

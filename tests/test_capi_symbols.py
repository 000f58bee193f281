"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/voronoids_b200.h declares.
No compute call is made here (this container has no GPU); error paths that need no device are exercised."""
import ctypes as C
import os
import re

import pytest

from voronoids_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_lib():
    from voronoids_b200 import build
    so = build.build()
    return _capi.bind(C.CDLL(so))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "voronoids_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vor_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_bindings_agree():
    assert sorted(_capi.SYMBOLS) == header_symbols()


def test_library_exports_every_declared_symbol(product_lib):
    for name in header_symbols():
        assert hasattr(product_lib, name), name


def test_argument_errors_need_no_device(product_lib):
    h = _capi.tree_p()
    assert product_lib.vor_tree_create_device(5, None, 10, 0, None, C.byref(h)) != 0   # bad dim
    assert product_lib.vor_tree_insert(None, None, 0, 1) == 10
    assert product_lib.vor_set_option(b"no_such_option", 1.0) == -1
    assert product_lib.vor_set_option(b"stats", 0.0) == 0


def test_no_cpu_fallback_without_gpu(product_lib):
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = _capi.tree_p()
    pts = np.random.default_rng(0).random((100, 3))
    st = product_lib.vor_tree_create(3, pts.ctypes.data_as(_capi.dp), 100, 0, C.byref(h))
    assert st == 4, "without a CUDA device the product must fail with VOR_ERR_CUDA, not fall back to a CPU path"
    assert b"cuda" in product_lib.vor_last_error().lower() or product_lib.vor_last_error()


def test_emulation_library_is_not_the_product():
    from voronoids_b200 import _lib
    assert _lib.SO_PATH.endswith("voronoids_b200/libvoronoids_b200.so")
    src = open(os.path.join(ROOT, "voronoids_b200", "_lib.py")).read() + open(os.path.join(ROOT, "voronoids_b200", "api.py")).read()
    assert "emu" not in src and "oracle" not in src

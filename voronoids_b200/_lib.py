"""Loader of the product library.  There is no CPU fallback: if libvoronoids_b200.so is missing the import of
the compute API fails loudly, and every compute call fails with VOR_ERR_CUDA when no CUDA device is present."""
import ctypes as C
import os

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("VOR_SO", os.path.join(_HERE, "libvoronoids_b200.so"))  # VOR_SO: tuning builds of the same sources
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -m voronoids_b200.build` (nvcc, sm_100a). "
                "voronoids_b200 has no CPU fallback.")
        _LIB = _capi.bind(C.CDLL(SO_PATH))
    return _LIB

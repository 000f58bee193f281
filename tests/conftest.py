import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: large configuration")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def emu_lib():
    """Kernel bodies + host loop compiled for the CPU (tests/emu); logic tests only, never the product path."""
    from voronoids_b200 import _capi
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["make", "-s", "-C", d], stdout=subprocess.DEVNULL)
    return _capi.bind(C.CDLL(os.path.join(d, "libvor_kernel_emu.so")))


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a CUDA device."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from voronoids_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return {c["name"]: c for c in json.load(f)["cases"]}

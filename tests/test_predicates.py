"""Exact predicates: the oracle's expansion arithmetic and the engine's multi-word integer path (compiled for the
CPU by tests/emu) against fractions.Fraction on adversarial inputs.  The GPU build of the same code is checked by
tests/test_gpu_predicates.py."""
import ctypes as C

import numpy as np
import pytest

import predcases as pc
from voronoids_b200 import _capi

KINDS = ["orient2d", "orient3d", "incircle", "insphere"]
KIND_ID = {"orient2d": 0, "orient3d": 1, "incircle": 2, "insphere": 3}


def emu_pred(lib, kind, rows):
    a = np.ascontiguousarray(rows, dtype=np.float64)
    out = np.zeros(a.shape[0], dtype=np.int32)
    ne = C.c_uint64()
    st = lib.vor_predicates(KIND_ID[kind], a.ctypes.data_as(_capi.dp), a.shape[0], out.ctypes.data_as(_capi.i32p), C.byref(ne), 0)
    return st, out, ne.value


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_predicates_vs_fractions(oracle, kind):
    rows = pc.adversarial(kind, 600, seed=1)
    want = np.array([pc.EXACT[kind](r) for r in rows])
    got_f = getattr(oracle, kind)(rows)
    got_e = getattr(oracle, kind)(rows, exact_only=True)
    assert np.array_equal(got_f, want)
    assert np.array_equal(got_e, want)
    assert (want == 0).sum() > 50  # the generator really produces exact degeneracies


@pytest.mark.parametrize("kind", KINDS)
def test_engine_predicates_vs_fractions(emu_lib, kind):
    rows = pc.adversarial(kind, 600, seed=2)
    want = np.array([pc.EXACT[kind](r) for r in rows])
    st, got, n_exact = emu_pred(emu_lib, kind, rows)
    assert st == 0
    assert np.array_equal(got, want)
    assert n_exact > 100  # the filter sent the degenerate rows to the integer path


@pytest.mark.parametrize("kind", KINDS)
def test_engine_predicates_wide_exponent_range(emu_lib, oracle, kind):
    rows = pc.wide_range(kind, 300, seed=3, span=120)
    want = np.array([pc.EXACT[kind](r) for r in rows])
    st, got, _ = emu_pred(emu_lib, kind, rows)
    assert st == 0
    assert np.array_equal(got, want)
    assert np.array_equal(getattr(oracle, kind)(rows), want)


def test_engine_predicates_range_error_is_loud(emu_lib):
    # coordinates spanning > 318 bits exceed the limb budget: the engine must say so, not return a sign
    row = np.array([[1e-200, 1.0, 1.0, 1.0, 1e-200, 1.0, 1.0, 1.0, 1e100, 3.0, 3.0, 3.0 + 1e-200]])
    # force the exact path: coplanar up to rounding
    row[0, 9:12] = row[0, 0:3] + row[0, 3:6] - row[0, 6:9] * 0
    st, got, _ = emu_pred(emu_lib, "orient3d", np.array([[1e-300, 0, 0, 0, 1e-300, 0, 1e300, 1e300, 0, 1e300, 1e300, 0]]))
    assert st in (0, 7)
    rows = np.array([[2.0 ** -400, 0.0, 0.0, 2.0 ** 100, 0.0, 2.0 ** -400, 0.0, 2.0 ** 100, 0.0, 0.0, 0.0, 0.0]])
    want = pc.orient3d_exact(rows[0])
    st, got, ne = emu_pred(emu_lib, "orient3d", rows)
    if ne:  # exact path taken: either a correct sign or a loud range error
        assert st == 7 or got[0] == want
    else:
        assert got[0] == want


def test_sign_conventions(oracle, emu_lib):
    tet = [1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1]  # reference KAT tet (tests/test_geometry.rs:6-11)
    o = oracle.orient3d([tet])[0]
    inside = [0.4, 0.4, 0.4]
    outside = [2.0, 2.0, 2.0]
    s_in = oracle.insphere([tet + inside])[0]
    s_out = oracle.insphere([tet + outside])[0]
    assert o != 0 and s_in == o and s_out == -o
    assert emu_pred(emu_lib, "insphere", [tet + inside])[1][0] == s_in
    assert oracle.orient2d([[0, 0, 1, 0, 0, 1]])[0] == 1
    assert oracle.incircle([[0, 0, 1, 0, 0, 1, 0.5, 0.5]])[0] == 1
    assert oracle.incircle([[0, 0, 1, 0, 0, 1, 1.0, 1.0]])[0] == 0  # on the circle: not in conflict (strict <)


@pytest.mark.parametrize("kind", ["orient3d", "incircle", "insphere"])
def test_engine_double_double_stage_on_jittered_lattice(emu_lib, kind):
    """dd_stage.cuh (compiled into the emulation build): near-degenerate lattice configurations at three coordinate
    scales get the exact sign, whether the double-double stage or the integers settle them."""
    dim, npts = {"orient3d": (3, 4), "incircle": (2, 4), "insphere": (3, 5)}[kind]
    rng = np.random.default_rng(11)
    for jit in (1e-9, 1e-13, 1e-16, 0.0):
        for scale in (1.0, 1e-40, 1e40):
            base = rng.integers(0, 4, size=(300, npts, dim)).astype(float) / 171.0
            rows = ((base + jit / 171.0 * rng.uniform(-1, 1, size=base.shape)) * scale).reshape(300, -1)
            want = np.array([pc.EXACT[kind](r) for r in rows])
            st, got, _ = emu_pred(emu_lib, kind, rows)
            assert st == 0 and np.array_equal(got, want), (kind, jit, scale)

"""The cached-circumsphere filter of the engine (sphere.cuh, compiled for the CPU by tests/emu) may only certify what the
exact predicate confirms -- on well-shaped, sliver, needle, tiny and nearly flat simplices, with queries aimed at the
sphere from 1e-3 down to 1e-16 relative distance, at vertices and at vertices nudged by a few ulps.  The GPU build of the
same code is checked by tests/test_gpu_engine.py::test_gpu_sphere_filter_is_certified."""
import numpy as np
import pytest

import spherecases as sc


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-3, 0.0), (1.0, 1e3), (1e6, -3e6)])
def test_sphere_filter_never_certifies_a_wrong_answer(emu_lib, oracle, dim, scale, offset):
    decided, verdict, want = sc.check_filter(emu_lib, oracle, dim, 20000, seed=5 + dim, scale=scale, offset=offset)
    # the filter is useful, not only safe: generic queries are decided
    assert decided > 0.1
    assert (verdict > 0).sum() > 100 and (verdict < 0).sum() > 100


@pytest.mark.parametrize("dim", [2, 3])
def test_sphere_filter_decides_generic_queries(emu_lib, oracle, dim):
    """random simplices in the unit box with random queries: everything but a ~1e-5 shell is decided"""
    rng = np.random.default_rng(1)
    n = 50000
    rows = rng.random((n, (dim + 2) * dim))
    verdict, _ = sc.sphere_filter(emu_lib, dim, np.full(dim, 0.5), 0.5 * dim, rows)
    want, flat = sc.exact_inside(oracle, dim, rows)
    assert not flat.any()
    assert np.all((verdict == 0) | (verdict == np.where(want > 0, 1, -1)))
    assert (verdict == 0).mean() < 2e-3


def test_sphere_filter_own_vertices_are_never_inside(emu_lib):
    rng = np.random.default_rng(2)
    for dim in (2, 3):
        simp = rng.random((4000, dim + 1, dim))
        for k in range(dim + 1):
            rows = np.concatenate([simp.reshape(4000, -1), simp[:, k, :]], axis=1)
            verdict, _ = sc.sphere_filter(emu_lib, dim, np.full(dim, 0.5), 0.5 * dim, rows)
            assert np.all(verdict == 0)   # on the sphere: inside the undecided shell, never certified either way

"""Kernel logic + host round loop of voronoids_b200/csrc, compiled for the CPU by tests/emu (sequential launches),
against the exact oracle.  This checks the algorithm in the GPU-less container; the product build of the very
same sources is checked on the B200 by tests/test_gpu_engine.py."""
import numpy as np
import pytest

import enginecases as ec
from voronoids_b200 import _capi, pointgen


@pytest.mark.parametrize("dim,kind,n", [(3, "uniform", 4000), (2, "uniform", 8000), (3, "clustered", 4000), (3, "lattice", 4000),
                                        (2, "clustered", 4000), (2, "lattice", 4000)])
def test_emu_matches_oracle(emu_lib, oracle, dim, kind, n):
    st = ec.check_against_oracle(emu_lib, oracle, pointgen.make(kind, n, dim, 3))
    assert st["winners"] == n


@pytest.mark.parametrize("dim", [2, 3])
def test_emu_tiny_inputs(emu_lib, oracle, dim):
    ec.case_tiny(emu_lib, oracle, dim)


@pytest.mark.parametrize("dim", [2, 3])
def test_emu_awkward_sizes_and_many_small_inserts(emu_lib, oracle, dim):
    ec.case_awkward_sizes(emu_lib, oracle, dim)


def test_emu_edge_wedge_ties(emu_lib, oracle):
    ec.case_edge_wedge_ties(emu_lib, oracle)


@pytest.mark.parametrize("dim", [2, 3])
def test_emu_incremental_insert(emu_lib, oracle, dim):
    ec.case_incremental(emu_lib, oracle, dim, 700, 5000)


@pytest.mark.parametrize("dim", [2, 3])
def test_emu_batch_of_sets(emu_lib, oracle, dim):
    ec.case_batch(emu_lib, oracle, dim, [1500, 300, 2000, 2, 777])


@pytest.mark.parametrize("dim", [2, 3])
def test_emu_duplicates_are_dropped_and_reported(emu_lib, oracle, dim):
    ec.case_duplicates(emu_lib, oracle, dim)


@pytest.mark.parametrize("dim", [2, 3])
def test_emu_locate(emu_lib, oracle, dim):
    ec.case_locate(emu_lib, oracle, dim)


@pytest.mark.parametrize("dim", [2, 3])
def test_emu_scheduler_matches_reference_restatement(emu_lib, oracle, dim):
    ec.case_scheduler(emu_lib, oracle, dim, n=1500, nq=300)


@pytest.mark.parametrize("dim", [2, 3])
def test_emu_export_vertices(emu_lib, dim):
    ec.case_export_vertices(emu_lib, dim, n=1200)


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0]])
def test_emu_batch_over_devices(emu_lib, oracle, devices):
    ec.case_batch_devices(emu_lib, oracle, 3, devices)


@pytest.mark.parametrize("dim", [2, 3])
def test_emu_check_delaunay_rejects_broken_meshes(emu_lib, dim):
    ec.case_check_delaunay_rejects(emu_lib, dim)


@pytest.mark.parametrize("dim,chunk_sets,chunk_points", [(3, 3, 0), (2, 2, 0), (3, 100, 2500)])
def test_emu_batch_stream(emu_lib, oracle, dim, chunk_sets, chunk_points):
    ec.case_batch_stream(emu_lib, oracle, dim, [900, 300, 1200, 2, 777, 1000, 450], chunk_sets, chunk_points)


def test_emu_overflow_scratch_and_compaction(emu_lib, oracle):
    # tiny regular slots force the overflow path; a small attempt budget forces many rounds + list compaction
    emu_lib.vor_set_option(b"capk", 8.0)
    emu_lib.vor_set_option(b"min_attempt", 512.0)
    try:
        st = ec.check_against_oracle(emu_lib, oracle, pointgen.uniform(20000, 3, 9))
        assert st["compactions"] > 0
    finally:
        emu_lib.vor_set_option(b"capk", 64.0)
        emu_lib.vor_set_option(b"min_attempt", 8192.0)


def test_emu_capacity_error_is_loud(emu_lib):
    emu_lib.vor_set_option(b"capk", 4.0)
    emu_lib.vor_set_option(b"big_capk", 8.0)
    try:
        with pytest.raises(_capi.VorError) as ei:
            _capi.Tree(emu_lib, pointgen.uniform(3000, 3, 1))
        assert ei.value.status == 6
    finally:
        emu_lib.vor_set_option(b"capk", 64.0)
        emu_lib.vor_set_option(b"big_capk", 8192.0)


def test_emu_point_outside_super_simplex_is_loud(emu_lib):
    pts = pointgen.uniform(200, 3, 1)
    t = _capi.Tree(emu_lib, pts, insert=False)
    with pytest.raises(_capi.VorError) as ei:
        t.insert(np.array([[1e6, 1e6, 1e6]]))
    assert ei.value.status == 8
    t.close()


def test_emu_owner_epoch_reset(emu_lib, oracle):
    # 2^17+ points in a stage leave 12 epoch bits; a tiny slot cap makes many rounds so the epoch wraps
    emu_lib.vor_set_option(b"slot_cap", 64.0)
    try:
        pts = pointgen.uniform(2500, 2, 4)
        st = ec.check_against_oracle(emu_lib, oracle, pts)
        assert st["rounds"] > 50
    finally:
        emu_lib.vor_set_option(b"slot_cap", float(1 << 19))


def test_cpp_header_program_on_the_emulation(emu_lib, oracle, tmp_path):
    """include/voronoids.hpp compiles and a program written against it (the crate's API names) gives the oracle's edges;
    linked to the kernel emulation here, to the product library in tests/test_gpu_engine.py"""
    import os
    import cppcase
    cppcase.run(tmp_path, oracle, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "libvor_kernel_emu.so"), n=1500)

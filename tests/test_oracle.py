"""The CPU oracle against everything the reference's own tests pin and against Qhull.

Known-answer material held by the reference:
  tests/test_geometry.rs:5-15       circumsphere KAT
  tests/test_delaunay_tree.rs:20    DelaunayTree::<3,4>::new(..).max_simplex_id == 4
  tests/test_delaunay_tree.rs:42-44 DelaunayTree::<2,3>::new(4 fixed points).max_simplex_id == 3
  tests/test_delaunay_tree.rs:23-37 100 sequential + 1000 parallel 3D inserts, check_delaunay (no panic)
  tests/test_delaunay_tree.rs:50-58 1000 sequential 2D inserts, check_delaunay
  tests/test_scheduler.rs:6-37      make_queue + find_placement on 1000 points over a 1000-point tree (no panic)
"""
import math

import numpy as np
import pytest

from voronoids_b200 import pointgen


def test_circumsphere_kat(oracle):
    c, r = oracle.ref_circumsphere([[1.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    assert c.tolist() == [0.5, 0.5, 0.5]
    assert r == 0.8660254037844386


def test_in_sphere_is_strict(oracle):
    assert oracle.ref_in_sphere([0.0, 0.0, 0.0], [0.5, 0.0, 0.0], 1.0)
    assert not oracle.ref_in_sphere([1.0, 0.0, 0.0], [0.0, 0.0, 0.0], 1.0)  # on the sphere: dist^2 < r^2 is false
    assert not oracle.ref_in_sphere([0.0, 0.0], [0.0, 0.0], 0.0)            # ghost simplices (radius 0) never conflict


def test_bounding_sphere_and_super_simplex(oracle):
    pts = pointgen.uniform(1000, 3, 0) * 2 - 1        # tests/test_geometry.rs:17-33 uses Uniform(-1,1)
    c, r = oracle.ref_bounding_sphere(pts)
    lo, hi = pts.min(0), pts.max(0)
    assert np.array_equal(c, (hi + lo) / 2.0)
    assert r == math.sqrt(sum((hi - c) ** 2)) or r == math.sqrt(sum((lo - c) ** 2))
    sup, c2, r10 = oracle.ref_super_simplex(pts)
    assert r10 == r * 10.0 and np.array_equal(c, c2)
    assert sup[0].tolist() == [c[0], c[1], c[2] + r10]
    assert sup[1].tolist() == [c[0] + r10, c[1], c[2] - r10]
    a1 = 2.0 * math.pi / 3.0
    assert sup[2].tolist() == [c[0] + r10 * math.cos(a1), c[1] + r10 * math.sin(a1), c[2] - r10]
    # a point sitting on a box corner fails the strict in_sphere test -> radius * 1.5 (geometry.rs:132-140)
    corner = np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.5, 0.2, 0.1]])
    c3, r3 = oracle.ref_bounding_sphere(corner)
    half = math.sqrt(0.75)
    assert r3 in (half, half * 1.5) and r3 >= half


def test_reference_structural_constants(oracle):
    pts = pointgen.uniform(1000, 3, 0)
    t = oracle.RefDelaunay(pts, mode="new")
    assert t.counts()["max_simplex_id"] == 4 and t.counts()["vertices"] == 8
    t2 = oracle.RefDelaunay(np.array([[0.3, 0.1], [1.0, 0.2], [0.1, 1.0], [0.5, 0.5]]), mode="new")
    assert t2.counts()["max_simplex_id"] == 3 and t2.counts()["vertices"] == 6


def test_reference_3d_sequence(oracle):
    """tests/test_delaunay_tree.rs:6-38 with our generator: 100 sequential, then 1000 through add_points_to_tree."""
    first = pointgen.uniform(1000, 3, 0)
    second = pointgen.uniform(1000, 3, 0, first=1000)
    allp = np.concatenate([first[:100], second])
    # the reference builds the super simplex from `vertices` (the first 1000), then inserts vertices[:100] and vertices2
    from oracle import oracle as O
    import ctypes as C
    L = O.lib()
    h = L.vo_ref_create(3, first.ctypes.data_as(C.POINTER(C.c_double)), 1000)
    a = np.ascontiguousarray(first[:100])
    assert L.vo_ref_insert_sequential(h, a.ctypes.data_as(C.POINTER(C.c_double)), 100, 0) == 0
    assert L.vo_ref_add_points_to_tree(h, second.ctypes.data_as(C.POINTER(C.c_double)), 1000, 100, 0) == 0
    out = (C.c_uint64 * 4)()
    L.vo_ref_counts(h, out)
    assert out[0] == 8 + 1100
    assert L.vo_ref_check_delaunay(h) == 1
    m = L.vo_ref_edges(h, None, 0)
    e = np.zeros((m, 2), dtype=np.uint32)
    L.vo_ref_edges(h, e.ctypes.data_as(C.POINTER(C.c_uint32)), m)
    L.vo_ref_destroy(h)
    sup = oracle.ref_super_simplex(first)[0]
    assert np.array_equal(e, oracle.ExactDelaunay(allp, super_vertices=sup).edges())


def test_reference_2d_sequence(oracle):
    """tests/test_delaunay_tree.rs:40-59: super simplex from 4 fixed points, 1000 sequential 2D inserts."""
    import ctypes as C
    L = oracle.lib()
    fixed = np.array([[0.3, 0.1], [1.0, 0.2], [0.1, 1.0], [0.5, 0.5]])
    pts = 0.1 + 0.8 * pointgen.uniform(1000, 2, 0)
    h = L.vo_ref_create(2, fixed.ctypes.data_as(C.POINTER(C.c_double)), 4)
    assert L.vo_ref_insert_sequential(h, pts.ctypes.data_as(C.POINTER(C.c_double)), 1000, 0) == 0
    assert L.vo_ref_check_delaunay(h) == 1
    m = L.vo_ref_edges(h, None, 0)
    e = np.zeros((m, 2), dtype=np.uint32)
    L.vo_ref_edges(h, e.ctypes.data_as(C.POINTER(C.c_uint32)), m)
    L.vo_ref_destroy(h)
    sup = oracle.ref_super_simplex(fixed)[0]
    assert np.array_equal(e, oracle.ExactDelaunay(pts, super_vertices=sup).edges())


def test_reference_scheduler(oracle):
    """tests/test_scheduler.rs:6-37: placement rounds are 1-based and respect footprint order."""
    pts = pointgen.uniform(2000, 3, 0)
    t = oracle.RefDelaunay(pts[:1000], mode="split", n_seq=1000)
    pl = t.placement(pts[1000:])
    assert pl.min() == 1 and pl.max() >= 2 and len(pl) == 1000


@pytest.mark.parametrize("dim,kind,n", [(3, "uniform", 3000), (2, "uniform", 5000), (3, "clustered", 3000)])
def test_exact_oracle_equals_float_restatement_and_qhull(oracle, dim, kind, n):
    from scipy.spatial import Delaunay
    pts = pointgen.make(kind, n, dim, 0)
    ex = oracle.ExactDelaunay(pts)
    assert ex.validate() == 0
    ref = oracle.RefDelaunay(pts)
    assert ref.err == 0
    assert np.array_equal(ref.edges(), ex.edges())
    sup = oracle.ref_super_simplex(pts)[0]
    q = Delaunay(np.vstack([sup, pts]))
    s1 = np.sort(q.simplices, axis=1)
    s1 = s1[np.lexsort(s1.T[::-1])]
    s2 = np.sort(ex.simplices(), axis=1)
    s2 = s2[np.lexsort(s2.T[::-1])]
    assert np.array_equal(s1, s2)


def test_oracle_matches_golden(oracle, golden):
    for name in ("u3_10k", "u2_10k"):
        g = golden[name]
        e = oracle.ExactDelaunay(pointgen.make(g["kind"], g["n"], g["dim"], g["seed"])).edges()
        assert len(e) == g["n_edges"] and oracle.edge_sha256(e) == g["sha256"]


def test_brute_force_empty_sphere(oracle):
    """check_delaunay as the reference does it (O(S*V), delaunay_tree.rs:512-541) on the exact oracle's output."""
    pts = pointgen.uniform(300, 3, 2)
    ex = oracle.ExactDelaunay(pts)
    sup = oracle.ref_super_simplex(pts)[0]
    allp = np.vstack([sup, pts])
    for s in ex.simplices():
        if s.min() < 4:
            continue
        rows = np.concatenate([np.tile(allp[s].reshape(-1), (len(allp), 1)), allp], axis=1)
        sign = oracle.insphere(rows) * oracle.orient3d([allp[s].reshape(-1)])[0]
        inside = np.nonzero(sign > 0)[0]
        assert len(inside) == 0


def test_float_restatement_on_near_degenerate_input_is_classified(oracle):
    """on the jittered lattice the reference's float circumsphere test may take wrong decisions
    (or panic).  The exact oracle is the authority there; the disagreement is measured, not hidden."""
    pts = pointgen.make("lattice", 2000, 3, 0)
    ex = oracle.ExactDelaunay(pts)
    assert ex.validate() == 0
    ref = oracle.RefDelaunay(pts)
    if ref.err == 0:
        a = set(map(tuple, ref.edges().tolist()))
        b = set(map(tuple, ex.edges().tolist()))
        # the float build is close to, but not necessarily equal to, the Delaunay graph
        assert len(a ^ b) <= 0.05 * len(b)
    else:
        assert ref.err in (1, 2, 3)  # the Rust original would have panicked

"""The Python surface of the reference (PyO3 module `voronoids`, /root/reference/src/lib.rs:12-134) as mirrored by
voronoids_b200.api -- logic tests on the CPU emulation of the kernels (tests/emu); tests/test_gpu_engine.py repeats the
essential ones on the device.  The emulation library is swapped in for the product library HERE ONLY (monkeypatch)."""
import time
from collections.abc import Mapping

import numpy as np
import pytest

from voronoids_b200 import pointgen


@pytest.fixture()
def vb(emu_lib, monkeypatch):
    import voronoids_b200
    from voronoids_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", emu_lib)
    return voronoids_b200


def test_import_voronoids_is_the_same_module_surface(vb):
    import voronoids                                       # lib.rs:127-134  #[pymodule] fn voronoids
    assert voronoids.delaunay is vb.delaunay
    assert voronoids.PyDelauanyTree is vb.PyDelauanyTree and voronoids.PySimplex is vb.PySimplex and voronoids.PyVertex is vb.PyVertex
    pts = pointgen.uniform(300, 3, 4)
    tree = voronoids.delaunay([list(map(float, p)) for p in pts])    # any sequence of 3-sequences, like Vec<[f64;3]>
    assert tree.max_simplex_id >= 4 and tree.check_delaunay()


@pytest.mark.parametrize("dim", [2, 3])
def test_simplices_and_vertices_are_lazy_mappings(vb, oracle, dim):
    n = 3000
    pts = pointgen.uniform(n, dim, 7)
    tree = vb.delaunay(pts)
    m = dim + 1
    S, V = tree.simplices, tree.vertices
    assert isinstance(S, Mapping) and isinstance(V, Mapping) and not isinstance(S, dict) and not isinstance(V, dict)
    v, nb, c, r = tree.simplex_arrays()
    assert len(S) == len(v) + m and len(V) == 2 * m + n
    first = m + 1
    # items agree with the exported arrays
    for k in (first, first + 1, first + len(v) // 2, first + len(v) - 1):
        s = S[k]
        i = k - first
        assert s.vertices == v[i].tolist() and s.center == c[i].tolist() and s.radius == r[i] and len(s.neighbors) == m
    # adjacency is symmetric through the mapping, ghost simplices included (delaunay_tree.rs:467-502)
    rng = np.random.default_rng(0)
    for k in rng.integers(first, first + len(v), size=200).tolist() + list(range(1, m + 1)):
        for j in S[k].neighbors:
            assert j in S and k in S[j].neighbors
    # ghosts: radius 0, centre 0, first vertex is a ghost vertex id
    for g in range(1, m + 1):
        assert S[g].radius == 0.0 and S[g].center == [0.0] * dim and S[g].vertices[0] == m + g - 1
    with pytest.raises(KeyError):
        S[first + len(v)]
    assert (first + len(v)) not in S and 0 not in S and "x" not in S
    assert list(S)[:m] == list(range(1, m + 1)) and sum(1 for _ in S) == len(S)
    # vertices: input point i has id 2M + i (delaunay_tree.rs:173-174); every incident simplex contains the vertex
    for i in (0, 1, n // 2, n - 1):
        pv = V[2 * m + i]
        assert pv.point == pts[i].tolist() and len(pv.simplex) > 0
        assert all((2 * m + i) in S[t].vertices for t in pv.simplex)
    sv = tree.super_simplex()[0]
    assert V[0].point == sv[0].tolist() and V[m].point == sv[0].tolist()      # ghost vertex = copy of a super vertex
    # the edge list derived from the lazy simplices equals the canonical one
    e = set()
    for k in range(first, first + len(v)):
        vs = [q - 2 * m for q in S[k].vertices if q >= 2 * m]
        e.update((min(a, b), max(a, b)) for a in vs for b in vs if a != b)
    assert sorted(e) == [tuple(x) for x in tree.edges().tolist()]
    assert np.array_equal(tree.edges(), oracle.ExactDelaunay(pts).edges())


def test_item_access_does_not_materialise_the_mesh(vb):
    pts = pointgen.uniform(20000, 3, 1)
    tree = vb.delaunay(pts)
    S = tree.simplices                      # exports the arrays once
    t0 = time.perf_counter()
    for k in range(5, 1005):
        S[k]
    dt = (time.perf_counter() - t0) / 1000
    assert dt < 1e-3, f"{dt * 1e6:.0f} us per item"
    assert tree.simplices is S              # cached until the tree changes
    tree.add_points_to_tree(pointgen.uniform(10, 3, 2) * 0.5 + 0.25)
    assert tree.simplices is not S


def test_empty_tree_views(vb):
    pts = pointgen.uniform(50, 3, 3)
    tree = vb.DelaunayTree.new(pts)         # DelaunayTree::new only (delaunay_tree.rs:390-510)
    assert tree.max_simplex_id == 4         # tests/test_delaunay_tree.rs:20
    S = tree.simplices
    assert sorted(S) == [0, 1, 2, 3, 4] and S[0].vertices == [0, 1, 2, 3] and S[0].neighbors == [1, 2, 3, 4]
    assert S[1].vertices == [4, 0, 1, 2] and S[1].neighbors == [0] and S[1].radius == 0.0
    assert len(tree.vertices) == 8

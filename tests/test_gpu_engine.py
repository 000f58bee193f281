"""Parity tests proper: the product library (hand-written sm_100a CUDA behind the C ABI) on a B200 against the
exact oracle, the committed golden vectors, and size-independent properties at BASELINE.json's full sizes."""
import os

import numpy as np
import pytest

import enginecases as ec
import predcases as pc
from voronoids_b200 import _capi, pointgen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dim,kind,n", [(3, "uniform", 10_000), (3, "uniform", 200_000), (2, "uniform", 200_000), (3, "clustered", 100_000),
                                        (3, "lattice", 100_000), (2, "clustered", 100_000), (2, "lattice", 100_000)])
def test_gpu_matches_oracle(gpu_lib, oracle, dim, kind, n):
    st = ec.check_against_oracle(gpu_lib, oracle, pointgen.make(kind, n, dim, 3))
    assert st["winners"] == n


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_tiny_inputs(gpu_lib, oracle, dim):
    ec.case_tiny(gpu_lib, oracle, dim)


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_awkward_sizes_and_many_small_inserts(gpu_lib, oracle, dim):
    ec.case_awkward_sizes(gpu_lib, oracle, dim)


def test_gpu_edge_wedge_ties(gpu_lib, oracle):
    ec.case_edge_wedge_ties(gpu_lib, oracle)


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_incremental_insert(gpu_lib, oracle, dim):
    # examples/parallel_insert.rs shape at 1/10 scale: 10k "sequential" + 100k "parallel"
    ec.case_incremental(gpu_lib, oracle, dim, 10_000, 100_000)


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_batch_of_sets(gpu_lib, oracle, dim):
    ec.case_batch(gpu_lib, oracle, dim, [20_000, 300, 50_000, 2, 7777, 10_000])


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_duplicates(gpu_lib, oracle, dim):
    ec.case_duplicates(gpu_lib, oracle, dim)


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_locate(gpu_lib, oracle, dim):
    ec.case_locate(gpu_lib, oracle, dim, n=20000, nq=60)


def test_gpu_voronoi_dual(gpu_lib, oracle):
    import voronoids_b200 as vb
    pts = pointgen.uniform(5000, 3, 0)
    tree = vb.delaunay(pts)
    centres, ridges = tree.voronoi()
    v, nb, c, r = tree.simplex_arrays()
    assert centres.shape == (len(v), 3) and ridges.shape[1] == 2
    # every interior facet gives exactly one ridge
    assert len(ridges) == int((nb >= 0).sum()) // 2
    # a Voronoi vertex is equidistant from the 4 vertices of its simplex (float circumsphere of the reference)
    allp = np.vstack([tree.super_simplex()[0], np.zeros((4, 3)), pts])
    d = np.linalg.norm(allp[v] - centres[:, None, :], axis=2)
    real = (v >= 8).all(axis=1)
    assert np.allclose(d[real], r[real, None], rtol=1e-6)
    ids = tree.locate(pts[0] * 0.5 + pts[1] * 0.5)
    assert ids == sorted(ids) and len(ids) > 0 and all(i in tree.simplices for i in ids)


def test_gpu_overflow_scratch(gpu_lib, oracle):
    gpu_lib.vor_set_option(b"capk", 8.0)
    gpu_lib.vor_set_option(b"min_attempt", 2048.0)
    try:
        st = ec.check_against_oracle(gpu_lib, oracle, pointgen.uniform(60_000, 3, 9))
        assert st["compactions"] > 0
    finally:
        gpu_lib.vor_set_option(b"capk", 64.0)
        gpu_lib.vor_set_option(b"min_attempt", 8192.0)


def test_gpu_errors_are_loud(gpu_lib):
    pts = pointgen.uniform(200, 3, 1)
    t = _capi.Tree(gpu_lib, pts, insert=False)
    with pytest.raises(_capi.VorError) as ei:
        t.insert(np.array([[1e6, 1e6, 1e6]]))
    assert ei.value.status == 8
    t.close()
    with pytest.raises(_capi.VorError) as ei:
        _capi.Tree(gpu_lib, pointgen.uniform(1, 3, 1))
    assert ei.value.status == 2


@pytest.mark.parametrize("name", ["u3_10k", "u3_100k", "u3_1m", "u2_10k", "u2_1m", "c3_100k", "l3_100k", "c3_500k", "l3_500k", "u3_set1000_100k",
                                  "c3_5m", "l3_5m"])   # the last two: BASELINE.json configs[3] at its stated size
def test_gpu_golden_vectors(gpu_lib, oracle, golden, name):
    g = golden[name]
    t = _capi.Tree(gpu_lib, pointgen.make(g["kind"], g["n"], g["dim"], g["seed"]))
    try:
        if g["n"] >= 5_000_000:
            ok, fails = t.check_delaunay()
            assert ok, fails
        e = t.edges()
        assert len(e) == g["n_edges"]
        assert oracle.edge_sha256(e) == g["sha256"]
        assert t.edge_checksum() == (g["n_edges"], g["checksum64"])
        assert t.counts()["simplices"] == g["live_simplices"]
    finally:
        t.close()


@pytest.mark.slow
def test_gpu_full_size_10m(gpu_lib, oracle, golden):
    """BASELINE.json configs[2]: 3D uniform 10M points, edge set bit-exact (hash of the canonical list) and the
    size-independent properties: valid orientation, symmetric adjacency, locally Delaunay everywhere."""
    g = golden.get("u3_10m")
    pts = pointgen.uniform(10_000_000, 3, 0)
    t = _capi.Tree(gpu_lib, pts)
    try:
        ok, fails = t.check_delaunay()
        assert ok, fails
        n, ck = t.edge_checksum()
        e = t.edges()
        assert len(e) == n and np.all(e[:, 0] < e[:, 1])
        k = (e[:, 0].astype(np.uint64) << np.uint64(32)) | e[:, 1]
        assert np.all(k[1:] > k[:-1]), "edge list must be sorted and unique"
        assert ck == _capi.edge_checksum_host(e)
        cnt = t.counts()
        # Euler-type count for a triangulated ball: live simplices relate to the vertex/edge counts of P u S
        assert cnt["vertices"] == 8 + 10_000_000
        if g is not None:
            assert n == g["n_edges"] and oracle.edge_sha256(e) == g["sha256"]
            assert cnt["simplices"] == g["live_simplices"]
    finally:
        t.close()


def test_gpu_batch_stream_small(gpu_lib, oracle):
    ec.case_batch_stream(gpu_lib, oracle, 3, [20_000, 300, 50_000, 2, 7777, 10_000, 30_000], chunk_sets=3)
    ec.case_batch_stream(gpu_lib, oracle, 2, [20_000, 300, 50_000, 2, 7777, 10_000, 30_000], chunk_sets=100, chunk_points=60_000)


@pytest.mark.slow
def test_gpu_batch_stream_1024_sets_of_100k(gpu_lib, oracle, golden):
    """BASELINE.json configs[4] at 1/8 of its set count on ONE GPU: 1,024 independent 3D sets x 100k points (102.4M points,
    8 chunks of 128 sets) through the streaming driver, device-resident input generated on the device with the same
    generator; a seeded sample of 16 sets + set 0 (golden u3_set1000_100k) against the exact oracle."""
    import torch
    import voronoids
    n_sets, n = 1024, 100_000
    d = pointgen.uniform_sets_torch(n_sets, n, 3, 1000)
    off = np.arange(n_sets + 1, dtype=np.int64) * n
    ne, ck = voronoids.delaunay_batch_stream(d, off)
    g = golden["u3_set1000_100k"]
    assert int(ne[0]) == g["n_edges"] and int(ck[0]) == g["checksum64"]
    rng = np.random.default_rng(2026)
    for s in sorted(rng.choice(n_sets, size=16, replace=False).tolist()):
        pts = pointgen.uniform(n, 3, 1000 + s)
        assert np.array_equal(d[s * n:(s + 1) * n].cpu().numpy(), pts)       # same generator on the device
        e = oracle.ExactDelaunay(pts).edges()
        assert int(ne[s]) == len(e) and int(ck[s]) == _capi.edge_checksum_host(e), s
    assert (ne > 700_000).all() and len(set(ck.tolist())) == n_sets
    # host input (copies of the chunks overlap the rounds): the first 256 sets give the same per-set results
    h = d[:256 * n].cpu().numpy()
    ne2, ck2 = voronoids.delaunay_batch_stream(h, off[:257])
    assert np.array_equal(ne2, ne[:256]) and np.array_equal(ck2, ck[:256])


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_check_delaunay_rejects_broken_meshes(gpu_lib, dim):
    ec.case_check_delaunay_rejects(gpu_lib, dim, n=50_000)


def test_gpu_cpp_header_program(gpu_lib, oracle, tmp_path):
    """a program written against include/voronoids.hpp (the C++ mirror of the Rust API), linked to the product library"""
    import cppcase
    cppcase.run(tmp_path, oracle, os.path.join(ROOT, "voronoids_b200", "libvoronoids_b200.so"), n=30_000)


def test_gpu_python_lazy_getters_and_device_input(gpu_lib, oracle, golden):
    """`import voronoids` (lib.rs:127-134) on a 1M-point tree: device-resident input through __cuda_array_interface__ and
    DLPack (no host copy of the coordinates), `.simplices` / `.vertices` as lazy mappings (an item in < 1 ms, nothing of
    size O(n) materialised as Python objects)."""
    import time
    import torch
    import voronoids
    pts = pointgen.uniform(1_000_000, 3, 0)
    d = torch.from_numpy(pts).cuda()
    tree = voronoids.delaunay(d)
    e = tree.edges()
    g = golden["u3_1m"]
    assert len(e) == g["n_edges"] and oracle.edge_sha256(e) == g["sha256"]
    S = tree.simplices
    assert not isinstance(S, dict) and len(S) == g["live_simplices"] + 4
    t0 = time.perf_counter()
    for k in range(5, 2005):
        S[k]
    assert (time.perf_counter() - t0) / 2000 < 1e-3
    k = 5 + len(S) // 2
    assert all(k in S[j].neighbors for j in S[k].neighbors)
    V = tree.vertices
    assert V[8 + 12345].point == pts[12345].tolist() and all((8 + 12345) in S[t].vertices for t in V[8 + 12345].simplex)
    tree.close()

    class OnlyDLPack:   # an array type that speaks DLPack but not __cuda_array_interface__
        def __init__(self, t): self.t = t
        def __dlpack__(self, stream=None): return self.t.__dlpack__()
        def __dlpack_device__(self): return self.t.__dlpack_device__()
    small = torch.from_numpy(pointgen.uniform(50_000, 2, 3)).cuda()
    t2 = voronoids.delaunay(OnlyDLPack(small))
    assert np.array_equal(t2.edges(), oracle.ExactDelaunay(small.cpu().numpy()).edges())
    with pytest.raises(ValueError):
        voronoids.delaunay(small.float())


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_sphere_filter_is_certified(gpu_lib, oracle, dim):
    """the cached-circumsphere filter (sphere.cuh) as compiled for the B200 never certifies what the exact predicate
    contradicts: adversarial rows (slivers, needles, near-sphere queries) at several scales and offsets"""
    import spherecases as sc
    for scale, offset in [(1.0, 0.0), (1e-3, 0.0), (1.0, 1e3), (1e6, -3e6)]:
        decided, verdict, want = sc.check_filter(gpu_lib, oracle, dim, 60_000, seed=11 + dim, scale=scale, offset=offset)
        assert decided > 0.1


@pytest.mark.parametrize("kind", ["orient2d", "orient3d", "incircle", "insphere"])
def test_gpu_predicates_vs_fractions(gpu_lib, oracle, kind):
    from voronoids_b200 import geometry
    rows = pc.adversarial(kind, 800, seed=5)
    want = np.array([pc.EXACT[kind](r) for r in rows])
    got, n_exact = geometry.predicate(kind, rows, return_exact_count=True)
    assert np.array_equal(got, want)
    assert n_exact > 100
    big = pc.adversarial(kind, 200_000, seed=6)
    assert np.array_equal(geometry.predicate(kind, big), getattr(oracle, kind)(big))
    wide = pc.wide_range(kind, 300, seed=7)
    assert np.array_equal(geometry.predicate(kind, wide), np.array([pc.EXACT[kind](r) for r in wide]))


def test_gpu_geometry_module(gpu_lib, oracle):
    from voronoids_b200 import geometry
    c, r = geometry.circumsphere([[1.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    assert c.tolist() == [0.5, 0.5, 0.5] and r == 0.8660254037844386      # tests/test_geometry.rs:5-15
    rng = np.random.default_rng(0)
    tets = rng.random((5000, 4, 3))
    cg, rg = geometry.circumsphere(tets)
    for i in range(0, 5000, 97):
        co, ro = oracle.ref_circumsphere(tets[i])
        assert np.array_equal(cg[i], co) and rg[i] == ro                  # same operation order, no FMA
    tris = rng.random((1000, 3, 2))
    cg, rg = geometry.circumsphere(tris)
    co, ro = oracle.ref_circumsphere(tris[5])
    assert np.array_equal(cg[5], co) and rg[5] == ro
    assert geometry.in_sphere([0.0, 0.0, 0.0], [0.5, 0.0, 0.0], 1.0) and not geometry.in_sphere([1.0, 0.0, 0.0], [0.0, 0.0, 0.0], 1.0)
    pts = pointgen.uniform(100_000, 3, 0) * 2 - 1
    cb, rb = geometry.bounding_sphere(pts)
    co, ro = oracle.ref_bounding_sphere(pts)
    assert np.array_equal(cb, co) and rb == ro


def test_gpu_python_api_matches_reference_shape(gpu_lib, oracle):
    """voronoids.delaunay(points) drop-in: attribute names and id conventions of lib.rs:12-134."""
    import voronoids_b200 as vb
    pts = pointgen.uniform(2000, 3, 0)
    tree = vb.delaunay(pts.tolist())           # the reference takes a list of [x,y,z]
    assert isinstance(tree, vb.PyDelauanyTree)
    verts, simps = tree.vertices, tree.simplices
    assert len(verts) == 8 + 2000 and verts[8].point == pts[0].tolist()
    assert tree.max_simplex_id >= max(simps)
    ex = oracle.ExactDelaunay(pts)
    real = {k: s for k, s in simps.items() if k > 4}
    assert len(real) == ex.stats()["live"]
    for k in (1, 2, 3, 4):
        assert simps[k].radius == 0.0 and simps[k].center == [0.0, 0.0, 0.0]  # ghosts, delaunay_tree.rs:467-502
    some = next(iter(real.values()))
    assert len(some.vertices) == 4 and len(some.center) == 3 and len(some.neighbors) == 4
    # adjacency is symmetric and Vertex.simplex lists are consistent with Simplex.vertices
    for k, s in list(real.items())[:500]:
        for nb in s.neighbors:
            assert k in simps[nb].neighbors
        for v in s.vertices:
            assert k in verts[v].simplex
    # circumspheres follow the reference's float formulas
    co, ro = oracle.ref_circumsphere(np.array([verts[v].point for v in some.vertices]))
    assert some.center == co.tolist() and some.radius == ro


def test_gpu_edges_pinned_block_outlives_tree(gpu_lib):
    """vor_tree_edges_host hands the caller a pinned block (zero-copy numpy view); same bytes as vor_tree_edges."""
    t = _capi.Tree(gpu_lib, pointgen.uniform(50_000, 3, 5))
    a = t.edges(pinned=True)
    b = t.edges(pinned=False)
    t.close()
    assert a.dtype == np.uint32 and a.shape == b.shape and np.array_equal(a, b)
    assert isinstance(a.base, _capi._HostBlock)
    ptr = a.base._ptr
    del a                                   # the block goes back to the library's pool ...
    assert gpu_lib.vor_host_free(ptr) != 0  # ... so a second free is refused loudly
    assert gpu_lib.vor_host_free(12345) != 0


ENGINE_DEFAULTS = {"red": 1, "commit_smem": 1, "split_exact": 1, "carry_frac": 0.125, "pdl": 131072, "mid_twin": 1}


@pytest.mark.parametrize("opts", [{"red": 0}, {"commit_smem": 0}, {"split_exact": 0}, {"split_exact": 0, "commit_smem": 0},
                                  {"carry_frac": 0.0}, {"carry_frac": 0.5}, {"pdl": 0}, {"carry_frac": 0.5, "pdl": 0, "mid_twin": 0}])
def test_gpu_engine_options_keep_parity(gpu_lib, oracle, opts):
    """every scheduling / layout option of the engine yields the oracle's edge set (3D and 2D)"""
    try:
        for k, v in opts.items():
            gpu_lib.vor_set_option(k.encode(), float(v))
        for dim in (3, 2):
            st = ec.check_against_oracle(gpu_lib, oracle, pointgen.uniform(120_000, dim, 11))
            assert st["winners"] == 120_000
    finally:
        for k, v in ENGINE_DEFAULTS.items():
            gpu_lib.vor_set_option(k.encode(), float(v))


@pytest.mark.parametrize("mid", [1, 0])
def test_gpu_lattice_through_the_mid_twin(gpu_lib, oracle, mid):
    """jittered lattice (a third of the conflict tests leave the float sphere filter): the hot kernel's twin with the FP64 determinant
    stage inside, the double-double stage and the exact twin behind it -- and the same input with the twin switched off"""
    try:
        gpu_lib.vor_set_option(b"mid_twin", float(mid))
        st = ec.check_against_oracle(gpu_lib, oracle, pointgen.make("lattice", 150_000, 3, 2))
        assert st["winners"] == 150_000
    finally:
        gpu_lib.vor_set_option(b"mid_twin", 1.0)


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_scheduler_matches_reference_restatement(gpu_lib, oracle, dim):
    ec.case_scheduler(gpu_lib, oracle, dim)


def test_gpu_python_scheduler_module(gpu_lib, oracle):
    """voronoids_b200.scheduler mirrors scheduler::{make_queue, find_placement} (tests/test_scheduler.rs:6-37)"""
    import voronoids_b200 as vb
    pts = pointgen.uniform(2000, 3, 0)
    tree = vb.delaunay(pts[:1000])
    queue = vb.scheduler.make_queue(pts[1000:], tree)
    assert len(queue) == 1000 and queue[5][0] == 5 and queue[5][1] == [float(x) for x in pts[1005]]
    assert all(s in tree.simplices for s in queue[5][2])
    placement = vb.scheduler.find_placement(queue)
    ref = oracle.RefDelaunay(pts[:1000], mode="split", n_seq=1000).placement(pts[1000:])
    assert placement == [int(x) for x in ref]


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_export_vertices(gpu_lib, dim):
    ec.case_export_vertices(gpu_lib, dim, n=20_000)


@pytest.mark.parametrize("dim", [2, 3])
def test_gpu_batch_over_devices(gpu_lib, oracle, dim):
    """vor_delaunay_batch with every visible device (and device 0 listed twice: two host threads on one GPU)"""
    import torch
    ec.case_batch_devices(gpu_lib, oracle, dim, list(range(torch.cuda.device_count())) + [0])

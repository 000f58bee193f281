"""Engine parity cases shared by the emulation tests (CPU, kernel logic) and the GPU tests (product library).
Every case goes through the C ABI (ctypes) and compares with the exact oracle."""
import numpy as np

from voronoids_b200 import _capi, pointgen


def canon_simplices(s):
    s = np.sort(np.asarray(s, dtype=np.int64), axis=1)
    return s[np.lexsort(s.T[::-1])]


def check_against_oracle(lib, O, pts, device=0, simplices=True):
    t = _capi.Tree(lib, pts, device=device)
    try:
        ok, fails = t.check_delaunay()
        assert ok, f"device mesh fails validation {fails}"
        ex = O.ExactDelaunay(pts)
        e = t.edges()
        eo = ex.edges()
        assert e.dtype == np.uint32 and e.shape == eo.shape, (e.shape, eo.shape)
        assert np.array_equal(e, eo), "edge list differs from the exact oracle"
        n, ck = t.edge_checksum()
        assert n == len(eo) and ck == _capi.edge_checksum_host(eo)
        sv, c, r = t.super_simplex()
        so, co, ro = O.ref_super_simplex(pts)
        assert np.array_equal(sv, so) and np.array_equal(c, co) and r == ro, "super simplex is not bit-identical"
        if simplices:
            m = pts.shape[1] + 1
            v, nb = t.simplices()
            # export ids: super k -> k, input i -> 2M + i ; oracle ids: super k -> k, input i -> M + i
            v2 = np.where(v >= 2 * m, v - m, v)
            assert np.array_equal(canon_simplices(v2), canon_simplices(ex.simplices())), "simplex set differs"
            cnt = t.counts()
            assert cnt["simplices"] == len(v) and cnt["vertices"] == 2 * m + len(pts)
        return t.stats()
    finally:
        t.close()


def case_incremental(lib, O, dim, n0, n1, seed=0):
    """DelaunayTree::new(all) then two insert calls (examples/parallel_insert.rs: 100k sequential + 1M parallel)."""
    pts = pointgen.uniform(n0 + n1, dim, seed)
    t = _capi.Tree(lib, pts, insert=False)
    try:
        t.insert(pts[:n0], mode=0)
        assert t.check_delaunay()[0]
        e0 = t.edges()
        # oracle for the first n0 points must use the super simplex of ALL points (tree was created from all)
        sup = O.ref_super_simplex(pts)[0]
        assert np.array_equal(e0, O.ExactDelaunay(pts[:n0], super_vertices=sup).edges())
        t.insert(pts[n0:], mode=1)
        assert t.check_delaunay()[0]
        assert np.array_equal(t.edges(), O.ExactDelaunay(pts).edges())
    finally:
        t.close()


def case_awkward_sizes(lib, O, dim):
    """sizes around the first stage (256), the stage hand-over thresholds and the attempt-slot floor (8192), and a tree grown by many
    small insert calls (each call starts new stages and must drain its last one)"""
    for n in (16, 100, 255, 256, 257, 300, 511, 513, 1000, 2049, 8191, 8193, 20_001):
        st = check_against_oracle(lib, O, pointgen.uniform(n, dim, 100 + n), simplices=n < 3000)
        assert st["winners"] == n
    pts = pointgen.uniform(6000, dim, 77)
    t = _capi.Tree(lib, pts, insert=False)
    try:
        cuts = [0, 1, 2, 5, 40, 41, 300, 301, 1500, 1501, 1502, 6000]
        for a, b in zip(cuts[:-1], cuts[1:]):
            t.insert(pts[a:b], mode=1 if (b - a) > 1 else 0)
        assert t.check_delaunay()[0]
        assert np.array_equal(t.edges(), O.ExactDelaunay(pts).edges())
    finally:
        t.close()


def case_edge_wedge_ties(lib, O):
    """3D edge list: the simplex whose wedge at the edge contains q = p_lo + dir emits the edge.  With dir chosen as the difference
    of two vertices of the mesh, q falls exactly ON faces and ON edge lines of real simplices: the tie rules (lower simplex id
    across the face / pivot around the edge) must still emit every edge exactly once.  Also the pivot-only path (the default:
    the wedge test is a recorded alternative, parity-green but no faster on the B200)."""
    # coordinates on a 2^-20 grid: differences and sums of points are exact, so q = p_lo + (p_b - p_a) IS a vertex / on an edge line
    pts = np.random.default_rng(21).integers(0, 1 << 20, size=(3000, 3)).astype(np.float64) / float(1 << 20)
    want = O.ExactDelaunay(pts).edges()
    dirs = [(0.0, 0.0, 0.0)]
    rng = np.random.default_rng(3)
    for _ in range(6):
        a, b = want[rng.integers(len(want))]
        dirs.append(tuple(pts[b] - pts[a]))
        dirs.append(tuple(pts[a] - pts[b]))
    dirs.append((1.0, 0.0, 0.0))
    try:
        for wedge in (1, 0):
            lib.vor_set_option(b"edge_wedge", float(wedge))
            for d in (dirs if wedge else dirs[:1]):
                for k, nm in enumerate((b"edge_dir_x", b"edge_dir_y", b"edge_dir_z")):
                    lib.vor_set_option(nm, float(d[k]))
                t = _capi.Tree(lib, pts)
                try:
                    z0 = t.stats()["exact_zero"]
                    assert np.array_equal(t.edges(), want), ("edge list differs", wedge, d)
                    if wedge and d != dirs[0] and d != dirs[-1]:
                        assert t.stats()["exact_zero"] > z0, "the direction was meant to provoke exact ties in the wedge test"
                finally:
                    t.close()
    finally:
        lib.vor_set_option(b"edge_wedge", 0.0)      # the default: pivots
        for nm in (b"edge_dir_x", b"edge_dir_y", b"edge_dir_z"):
            lib.vor_set_option(nm, 0.0)


def case_batch(lib, O, dim, sizes, seed0=1000):
    sets = [pointgen.uniform(n, dim, seed0 + s) for s, n in enumerate(sizes)]
    off = np.zeros(len(sets) + 1, dtype=np.int64)
    off[1:] = np.cumsum(sizes)
    allp = np.concatenate(sets, axis=0)
    t = _capi.Tree(lib, allp, set_offsets=off)
    try:
        assert t.check_delaunay()[0]
        e = t.edges()
        for s, p in enumerate(sets):
            a = np.searchsorted(e[:, 0], off[s], side="left")
            b = np.searchsorted(e[:, 0], off[s + 1], side="left")
            es = (e[a:b].astype(np.int64) - off[s]).astype(np.uint32)
            assert np.array_equal(es, O.ExactDelaunay(p).edges()), f"set {s} differs from its own triangulation"
            sv = t.super_simplex(s)[0]
            assert np.array_equal(sv, O.ref_super_simplex(p)[0])
    finally:
        t.close()


def case_duplicates(lib, O, dim):
    pts = pointgen.uniform(3000, dim, 5)
    dup = np.concatenate([pts, pts[:7]], axis=0)  # 7 exact duplicates appended
    t = _capi.Tree(lib, dup)
    try:
        assert t.duplicates, "duplicates must be reported (VOR_ERR_DUPLICATE_POINT)"
        assert t.check_delaunay()[0]
        e = t.edges()
        # the surviving copy of a duplicated pair may be either index; compare through coordinates
        eo = O.ExactDelaunay(pts).edges()
        rep = np.arange(len(dup))
        rep[len(pts):] = np.arange(7)
        ee = np.sort(rep[e.astype(np.int64)], axis=1)
        ee = np.unique(ee, axis=0)
        assert np.array_equal(ee.astype(np.uint32), eo)
        assert t.stats()["duplicates"] == 7
    finally:
        t.close()


def case_tiny(lib, O, dim):
    for n in (2, 3, 4, 5, 9):
        pts = pointgen.uniform(n, dim, 11 + n)
        check_against_oracle(lib, O, pts)
    # one point: zero-radius bounding sphere -> the reference panics ("No simplex found"); here a status code
    import pytest
    with pytest.raises(_capi.VorError) as ei:
        _capi.Tree(lib, pointgen.uniform(1, dim, 3))
    assert ei.value.status == 2
    # the reference's own 2D test points (tests/test_delaunay_tree.rs:42)
    if dim == 2:
        check_against_oracle(lib, O, np.array([[0.3, 0.1], [1.0, 0.2], [0.1, 1.0], [0.5, 0.5]]))


def case_locate(lib, O, dim, n=3000, nq=40):
    """DelaunayTree::locate (delaunay_tree.rs:33-58): the flooded conflict region equals the brute-force set of
    simplices whose open circumsphere contains the query (exact predicates), for points in general position."""
    pts = pointgen.uniform(n, dim, 0)
    t = _capi.Tree(lib, pts)
    try:
        v, nb = t.simplices()
        q = pointgen.uniform(nq, dim, 99)
        regs = t.locate(q)
        m = dim + 1
        allp = np.vstack([t.super_simplex()[0], np.zeros((m, dim)), pts])   # export ids: super, ghosts, 2M + i
        flat = allp[v].reshape(len(v), -1)
        orient = O.orient3d(flat) if dim == 3 else O.orient2d(flat)
        for qi in range(nq):
            rows = np.concatenate([flat, np.tile(q[qi], (len(v), 1))], axis=1)
            s = (O.insphere(rows) if dim == 3 else O.incircle(rows)) * orient
            assert np.array_equal(np.nonzero(s > 0)[0], regs[qi])
        # a query that coincides with an inserted vertex has an empty (strict) conflict region
        assert len(t.locate(pts[:1])[0]) == 0
    finally:
        t.close()


def case_scheduler(lib, O, dim, n=3000, nq=500):
    """scheduler::make_queue + find_placement (scheduler.rs:6-55, tests/test_scheduler.rs:6-37) against the restated
    reference: same 1-based rounds for the same queue, footprints sorted / unique / inside the export index space."""
    pts = pointgen.uniform(n, dim, 7)
    t = _capi.Tree(lib, pts)
    ref = O.RefDelaunay(pts, mode="split", n_seq=n)
    try:
        nlive = t.counts()["simplices"]
        for q in (0.3 + 0.4 * pointgen.uniform(nq, dim, 8), pointgen.uniform(nq, dim, 9)):
            off, ids = t.make_queue(q)
            assert off[0] == 0 and len(off) == nq + 1 and off[-1] == len(ids)
            regions = t.locate(q)
            for i in range(0, nq, 37):
                fp = ids[off[i]:off[i + 1]]
                assert len(fp) > 0 and np.all(np.diff(fp) > 0) and fp[0] >= 0 and fp[-1] < nlive
                assert set(regions[i].tolist()) <= set(fp.tolist())     # the conflict region lies inside its own 2-ring
            rounds = _capi.find_placement(lib, off, ids)
            assert rounds.min() == 1 and np.array_equal(rounds, ref.placement(q))
        # one entry, and an empty footprint is the reference's panic
        assert _capi.find_placement(lib, np.array([0, 3]), np.array([5, 6, 9]))[0] == 1
        try:
            _capi.find_placement(lib, np.array([0, 0]), np.zeros(0, dtype=np.int32))
            raise AssertionError("empty footprint accepted")
        except _capi.VorError as e:
            assert e.status == 1
    finally:
        t.close()


def case_export_vertices(lib, dim, n=2000):
    """vor_tree_export_vertices: reference id order, coordinates, Vertex.simplex as sorted CSR of export indices."""
    pts = pointgen.uniform(n, dim, 7)
    t = _capi.Tree(lib, pts)
    try:
        m = dim + 1
        coords, off, simps = t.vertices()
        v, _ = t.simplices()
        sv = t.super_simplex()[0]
        ghost_of = {3: [0, 0, 0, 1], 2: [0, 1, 2]}[dim]
        assert coords.shape == (2 * m + n, dim) and np.array_equal(coords[2 * m:], pts)
        assert np.array_equal(coords[:m], sv) and np.array_equal(coords[m:2 * m], sv[ghost_of])
        assert off[0] == 0 and off[-1] == len(simps) == len(v) * m
        inc = {}
        for i, row in enumerate(v.tolist()):
            for q in row:
                inc.setdefault(q, []).append(i)
        for q in range(len(coords)):
            assert simps[off[q]:off[q + 1]].tolist() == sorted(inc.get(q, []))
    finally:
        t.close()


def case_batch_devices(lib, O, dim, devices, sizes=(1500, 300, 2000, 2, 777, 1200, 64)):
    """vor_delaunay_batch (E1): every set's edge list equals the oracle's whichever device block it landed in."""
    sets = [pointgen.uniform(n, dim, 1000 + s) for s, n in enumerate(sizes)]
    off = np.concatenate([[0], np.cumsum([len(x) for x in sets])]).astype(np.int64)
    trees, shard = _capi.delaunay_batch_devices(lib, np.concatenate(sets, axis=0), off, devices)
    try:
        per = -(-len(sets) // len(devices))
        assert shard.tolist() == [min(len(sets), d * per) for d in range(len(devices) + 1)]
        for d, t in enumerate(trees):
            lo, hi = int(shard[d]), int(shard[d + 1])
            if hi == lo:
                assert t is None
                continue
            e = t.edges()
            base = off[lo]
            for s in range(lo, hi):
                a, b = off[s] - base, off[s + 1] - base
                i0, i1 = np.searchsorted(e[:, 0], a, side="left"), np.searchsorted(e[:, 0], b, side="left")
                got = (e[i0:i1] - np.uint32(a)).astype(np.uint32)
                assert np.array_equal(got, O.ExactDelaunay(sets[s]).edges()), (d, s)
    finally:
        for t in trees:
            if t is not None:
                t.close()


def case_check_delaunay_rejects(lib, dim, n=4000):
    """check_delaunay (delaunay_tree.rs:512-541) must REJECT a damaged mesh: each of the six failure classes of the
    checker is provoked through the test hook vor_debug_corrupt and must fire (the reference's own test discards the
    checker's verdict, tests/test_delaunay_tree.rs:37, and never feeds it a broken mesh)."""
    names = ["orientation", "dead neighbour", "asymmetric adjacency", "facet mismatch", "not Delaunay", "sphere filter"]
    for kind in range(6):
        t = _capi.Tree(lib, pointgen.uniform(n, dim, 21))
        try:
            ok, fails = t.check_delaunay()
            assert ok and not fails.any()
            assert lib.vor_debug_corrupt(t._h, kind) == 0
            ok, fails = t.check_delaunay()
            assert not ok, f"a mesh with a {names[kind]} defect was accepted"
            assert fails[kind] > 0, (names[kind], fails)
        finally:
            t.close()
    t = _capi.Tree(lib, pointgen.uniform(100, dim, 21))
    assert lib.vor_debug_corrupt(t._h, 9) == 10     # VOR_ERR_ARG
    t.close()


def case_batch_stream(lib, O, dim, sizes, chunk_sets, chunk_points=0, sample=None, first_seed=1000):
    """vor_delaunay_batch_stream: per-set (n_edges, checksum64) equal the exact oracle's for the sets in `sample`
    (default: all), whatever the chunking; the per-chunk callback hands out the chunk's canonical edge list."""
    sets = [pointgen.uniform(n, dim, first_seed + s) for s, n in enumerate(sizes)]
    off = np.zeros(len(sets) + 1, dtype=np.int64)
    off[1:] = np.cumsum(sizes)
    allp = np.concatenate(sets, axis=0)
    got = {}

    def on_chunk(first_set, n_sets, first_point, edges):
        assert first_point == off[first_set]
        for s in range(first_set, first_set + n_sets):
            lo, hi = off[s] - first_point, off[s + 1] - first_point
            a = np.searchsorted(edges[:, 0], lo, side="left")
            b = np.searchsorted(edges[:, 0], hi, side="left")
            got[s] = (edges[a:b].astype(np.int64) - lo).astype(np.uint32)
    ne, ck = _capi.delaunay_batch_stream(lib, allp, off, chunk_sets=chunk_sets, chunk_points=chunk_points, on_chunk=on_chunk)
    assert sorted(got) == list(range(len(sets)))
    for s in (range(len(sets)) if sample is None else sample):
        e = O.ExactDelaunay(sets[s]).edges()
        assert int(ne[s]) == len(e) and int(ck[s]) == _capi.edge_checksum_host(e), s
        assert np.array_equal(got[s], e)
    # without a callback, and with everything in one chunk: same per-set results
    ne2, ck2 = _capi.delaunay_batch_stream(lib, allp, off, chunk_sets=len(sets) + 5)
    assert np.array_equal(ne, ne2) and np.array_equal(ck, ck2)
    return ne, ck

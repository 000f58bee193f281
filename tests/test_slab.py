"""Slab mode: ONE triangulation over several ranks, gloo on the CPU with the kernel emulation standing in
for the product library on every rank.  The sorted union of the ranks' edge parts must be byte-identical to the canonical
edge list of the whole set (exact oracle = what a single-GPU run produces, tests/test_emu_engine.py)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, dim, kind, n, seed, coarse_div):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from voronoids_b200 import _capi, pointgen, slab
    lib = _capi.bind(C.CDLL(os.path.join(ROOT, "tests", "emu", "libvor_kernel_emu.so")))
    allp = torch.from_numpy(pointgen.make(kind, n, dim, seed))
    mine, gidx = slab.partition_by_axis(allp, world, rank, axis=0)
    res = slab.delaunay_slab(lib, mine, gidx, device=0, axis=0, coarse_div=coarse_div)
    full = slab.gather_edges(res.edges)
    infos = [None] * world
    dist.all_gather_object(infos, res.info)
    if rank == 0:
        q.put((full.tobytes(), full.shape, infos))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dim,kind,n,world,coarse_div", [(3, "uniform", 6000, 2, 16), (2, "uniform", 9000, 3, 16), (3, "clustered", 5000, 2, 8),
                                                         (3, "uniform", 4000, 4, 4), (2, "lattice", 4000, 2, 16)])
def test_slab_union_equals_single_triangulation(emu_lib, oracle, dim, kind, n, world, coarse_div):
    from voronoids_b200 import pointgen
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() * 7 + n + world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, dim, kind, n, 5, coarse_div)) for r in range(world)]
    for p in procs:
        p.start()
    raw, shape, infos = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = np.frombuffer(raw, dtype=np.uint32).reshape(shape)
    want = oracle.ExactDelaunay(pointgen.make(kind, n, dim, 5)).edges()
    assert got.shape == want.shape and got.tobytes() == want.tobytes(), "slab union differs from the single triangulation"
    # the decomposition is real: no rank held the whole set, halos were exchanged, every rank certified itself
    assert sum(i["own_points"] for i in infos) == n
    assert all(i["rounds"] >= 1 for i in infos)
    if world == 2 and kind == "uniform":
        assert all(i["tree_points"] < 0.85 * n for i in infos), infos
        assert all(i["halo_rows_received"] > 0 for i in infos)


def test_slab_single_rank_is_the_plain_path(emu_lib, oracle):
    import torch
    from voronoids_b200 import pointgen, slab
    pts = pointgen.uniform(3000, 3, 9)
    res = slab.delaunay_slab(emu_lib, torch.from_numpy(pts), torch.arange(3000), device=0)
    assert np.array_equal(res.edges, oracle.ExactDelaunay(pts).edges())


def _points_in_spheres_case(lib, device_points):
    """vor_points_in_spheres against fractions.Fraction: exact in-sphere counts of random points (some within ulps of the sphere)
    against a well-shaped simplex, a sliver and a nearly flat hull-like simplex with a super-vertex-sized edge."""
    from fractions import Fraction as F
    from voronoids_b200 import _capi
    rng = np.random.default_rng(7)
    simp = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]],
                     [[0.1, 0.1, 0.1], [0.9, 0.12, 0.1], [0.5, 0.8, 0.1000001], [0.45, 0.3, 0.0999999]],
                     [[0.2, 0.2, 0.999], [0.7, 0.3, 0.9991], [0.4, 0.8, 0.99905], [17.0, -3.0, 25.0]]], dtype=np.float64)
    # orientation as the mesh stores it: positive (swap two vertices where it is not)
    def det3(a, b, c, d):
        m = [[F(a[i]) - F(d[i]) for i in range(3)], [F(b[i]) - F(d[i]) for i in range(3)], [F(c[i]) - F(d[i]) for i in range(3)]]
        return (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
                + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]))
    pts = rng.random((4000, 3))
    pts[:8] = simp[0][[0, 1, 2, 3, 0, 1, 2, 3]]            # ON the first sphere: not strictly inside
    pts[8] = [0.5, 0.5, 0.5]
    pts[9] = np.nextafter(1.0, 2.0), 0.0, 0.0               # an ulp outside a vertex
    def insphere_exact(s, p):
        rows = []
        for v in s:
            d = [F(v[i]) - F(p[i]) for i in range(3)]
            rows.append(d + [d[0] * d[0] + d[1] * d[1] + d[2] * d[2]])
        def minor(r, cols):
            a, b, c = [[r[i][j] for j in cols] for i in range(3)]
            return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0])
        det = 0
        for i in range(4):
            others = [rows[j] for j in range(4) if j != i]
            det += (-1) ** (i + 3) * rows[i][3] * minor(others, [0, 1, 2])
        return det
    want = []
    for k in range(len(simp)):
        o = det3(*simp[k])
        assert o != 0
        if o < 0:
            simp[k][[0, 1]] = simp[k][[1, 0]]
            o = -o
        # sign convention probed with the centroid of the simplex (strictly inside its circumsphere)
        cen = simp[k].mean(axis=0)
        sgn = 1 if insphere_exact(simp[k], cen) > 0 else -1
        want.append(sum(1 for p in pts if sgn * insphere_exact(simp[k], p) > 0))
    inside = np.zeros(len(simp), dtype=np.uint64)
    st = lib.vor_points_in_spheres(3, C.c_void_p(device_points(pts)), len(pts), np.ascontiguousarray(simp).ctypes.data_as(_capi.dp), len(simp), 0,
                                   inside.ctypes.data_as(_capi.u64p))
    assert st == 0
    assert [int(x) for x in inside] == want, (inside, want)
    assert want[0] > 0 and want[0] < len(pts)


def test_points_in_spheres_is_exact(emu_lib):
    keep = []
    def host(pts):
        keep.append(np.ascontiguousarray(pts))
        return keep[-1].ctypes.data
    _points_in_spheres_case(emu_lib, host)


@pytest.mark.parametrize("dim", [3, 2])
def test_uncertified_list_matches_certify(emu_lib, dim):
    """vor_tree_uncertified_slab = vor_tree_certify_slab + the uncertified simplices themselves: same count and need, the list capped
    at `cap`, every listed simplex really reaches beyond the range"""
    from voronoids_b200 import _capi, pointgen
    pts = pointgen.uniform(2000, dim, 5)
    t = _capi.Tree(emu_lib, pts)
    try:
        own = np.ones(len(pts), dtype=np.uint8)
        ownp = own.ctypes.data_as(C.POINTER(C.c_uint8))
        n2, need2 = C.c_uint64(0), np.zeros(2)
        assert emu_lib.vor_tree_certify_slab(t._h, ownp, len(own), 0, 0.4, 0.6, 0.0, C.byref(n2), need2.ctypes.data_as(_capi.dp)) == 0
        assert n2.value > 0
        for cap in (1, 7, 4096):
            verts, reach, n, need = np.zeros((cap, dim + 1, dim)), np.zeros((cap, 2)), C.c_uint64(0), np.zeros(2)
            assert emu_lib.vor_tree_uncertified_slab(t._h, ownp, len(own), 0, 0.4, 0.6, 0.0, verts.ctypes.data_as(_capi.dp),
                                                     reach.ctypes.data_as(_capi.dp), cap, C.byref(n), need.ctypes.data_as(_capi.dp)) == 0
            assert n.value == n2.value and np.array_equal(need, need2)
            k = min(cap, n.value)
            assert np.all((reach[:k, 0] < 0.4) | (reach[:k, 1] > 0.6))
            # the listed coordinates are vertices of the tree (points of the set or super vertices)
            real = verts[:k].reshape(-1, dim)
            inset = (real[:, None, :] == pts[None, :64, :]).all(-1).any(-1) if k * (dim + 1) < 200 else None
            assert np.isfinite(real).all() and (inset is None or inset.dtype == bool)
        assert emu_lib.vor_tree_uncertified_slab(t._h, ownp, len(own), 0, 0.4, 0.6, 0.0, None, None, 0, C.byref(n2), need2.ctypes.data_as(_capi.dp)) != 0
    finally:
        t.close()

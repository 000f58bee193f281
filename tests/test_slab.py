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

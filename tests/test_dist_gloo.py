"""world_size-2 gloo test of the N>1 path: sets are sharded across ranks with no data-path collective; the barrier,
the max-over-ranks timing and the per-set gather are the only exchanges.  Each rank runs the kernel-logic emulation
(CPU) in place of the product library; every set must equal its own single-set triangulation (oracle)."""
import ctypes as C
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_SETS = 5
SIZES = [600, 900, 300, 1200, 450]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from voronoids_b200 import _capi, pointgen, sharding
    lib = _capi.bind(C.CDLL(os.path.join(ROOT, "tests", "emu", "libvor_kernel_emu.so")))
    mine = sharding.shard_range(N_SETS, world, rank)
    sets = [pointgen.uniform(SIZES[s], 3, 1000 + s) for s in mine]
    dist.barrier()
    local, tree = sharding.triangulate_sets(lib, sets)
    tree.close()
    res = sharding.gather_per_set({mine[k]: v for k, v in local.items()}, N_SETS)
    # the streaming driver on this rank's block (chunks of 2 sets): bench.py --workload b3_8192x100k does exactly this
    off = np.zeros(len(sets) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in sets])
    ne, ck = _capi.delaunay_batch_stream(lib, np.concatenate(sets, axis=0), off, chunk_sets=2)
    res_stream = sharding.gather_per_set({mine[k]: (int(ne[k]), int(ck[k])) for k in range(len(sets))}, N_SETS)
    assert res_stream == res
    tmax = sharding.max_over_ranks(1.0 + rank)
    tot = sharding.sum_over_ranks(sum(SIZES[s] for s in mine))
    if rank == 0:
        q.put((res, tmax, tot, list(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(emu_lib, oracle):
    from voronoids_b200 import _capi, pointgen, sharding
    assert list(sharding.shard_range(5, 2, 0)) == [0, 1, 2] and list(sharding.shard_range(5, 2, 1)) == [3, 4]
    assert list(sharding.shard_range(8192, 8, 7)) == list(range(7168, 8192))
    assert list(sharding.shard_range(3, 8, 5)) == []
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, tmax, tot, mine0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0 and tot == float(sum(SIZES)) and mine0 == [0, 1, 2]
    for s in range(N_SETS):
        e = oracle.ExactDelaunay(pointgen.uniform(SIZES[s], 3, 1000 + s)).edges()
        assert res[s] == (len(e), _capi.edge_checksum_host(e)), f"set {s}: sharded result differs from the single-set result"

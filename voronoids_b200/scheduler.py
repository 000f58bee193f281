"""Mirror of the reference's `pub mod scheduler` (/root/reference/src/scheduler.rs) on the device.

    make_queue(vertices, tree) -> [(id, vertex, neighbors), ...]     scheduler.rs:6-28
    find_placement(queue)      -> [round, ...] (1-based)             scheduler.rs:30-55

The engine itself does not schedule with these (rounds are resolved by atomicMin reservation on the device, see
DESIGN.md §3); they are here so that callers of the reference's scheduler find the same two functions.  `neighbors`
holds the ids used as keys of `tree.simplices`.  The reference's ghost simplices do not exist in the device store, so
footprints (and hence rounds) can differ from the reference's for points whose 2-ring reaches the super simplex's hull.
"""
import numpy as np

from . import _capi
from ._lib import lib


def make_queue(vertices, tree):
    v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, tree.dim)
    off, ids = tree._t.make_queue(v)
    first = tree.dim + 2   # id of export index 0 (after the reference's ghost simplices)
    return [(i, [float(x) for x in v[i]], (ids[off[i]:off[i + 1]] + first).astype(np.int64).tolist()) for i in range(v.shape[0])]


def find_placement(queue, device=0):
    off = np.zeros(len(queue) + 1, dtype=np.int64)
    for k, (_, _, nb) in enumerate(queue):
        off[k + 1] = off[k] + len(nb)
    ids = np.fromiter((s for _, _, nb in queue for s in nb), dtype=np.int64, count=int(off[-1]))
    rounds = _capi.find_placement(lib(), off, ids.astype(np.int32), device)
    # the reference indexes `placement` by the id stored in the queue entry (scheduler.rs:46)
    out = [0] * len(queue)
    for k, (i, _, _) in enumerate(queue):
        out[i] = int(rounds[k])
    return out

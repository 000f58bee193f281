"""Multi-GPU plumbing for the only place the path shards: batches of independent point sets.

One process per GPU (torch.distributed: nccl on the B200 box, gloo in CPU tests).  Sets are independent, so there is
NO data-path collective: each rank triangulates a contiguous block of sets; the only exchanges are the barrier, the
max-over-ranks of the device time, and a gather of per-set results (edge counts / checksums).
A single triangulation runs on one GPU ("replicas only" across GPUs).
"""
import torch
import torch.distributed as dist


def shard_range(n_units, world, rank):
    """Contiguous block of ceil(n_units / world) units for `rank`."""
    per = -(-n_units // world)
    lo = min(rank * per, n_units)
    hi = min(lo + per, n_units)
    return range(lo, hi)


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def max_over_ranks(x):
    """Time of a multi-GPU step = the slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_per_set(local, n_units):
    """local: {set index: result} of this rank -> list of n_units results on every rank (host-side gather)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local.get(i) for i in range(n_units)]
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged.get(i) for i in range(n_units)]


def triangulate_sets(lib, sets, device=0):
    """Triangulate this rank's sets in ONE device store; returns {local index: (n_edges, checksum64)} plus the tree."""
    import numpy as np
    from . import _capi
    off = np.zeros(len(sets) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in sets])
    tree = _capi.Tree(lib, np.concatenate(sets, axis=0), device=device, set_offsets=off)
    e = tree.edges()
    out = {}
    for s in range(len(sets)):
        a = np.searchsorted(e[:, 0], off[s], side="left")
        b = np.searchsorted(e[:, 0], off[s + 1], side="left")
        es = (e[a:b].astype(np.int64) - off[s]).astype(np.uint32)
        out[s] = (int(b - a), _capi.edge_checksum_host(es))
    return out, tree

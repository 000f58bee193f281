"""Slab mode: ONE Delaunay triangulation spread over several GPUs (north_star: "optional spatial-slab
decomposition with halo exchange over NVLink P2P; the slab result must match the single-GPU result").

The reference has one shared tree and one `add_points_to_tree` over it (/root/reference/src/delaunay_tree.rs:336-386); there is
nothing to port.  The construction here, one process per GPU (torch.distributed: NCCL send/recv over NVLink on the B200 box,
gloo in the CPU tests):

  1. bounds     every rank reduces its local box and its count of points outside the half-diagonal sphere (min / max / sum):
                all ranks build the SAME super simplex a single-GPU run would build (delaunay_tree.rs:392-406).
  2. coarse     a pseudo-random 1/coarse_div sample of the points (hash of the global index) is all-gathered and inserted
                by EVERY rank: it bounds the size of every circumsphere everywhere, hull included.
  3. slab       rank k owns the points whose `axis` coordinate lies in its range; it inserts its remaining points and a
                halo: the other ranks' points within a margin of its range, exchanged peer to peer.
  4. certify    vor_tree_certify_slab: every simplex around an owned point must have (circumsphere n data box) inside the
                range in which the tree holds every global point.  Simplices that reach further say how far; the halo
                is widened to that extent and the new points are inserted incrementally, until every rank is certified.
  5. edges      a certified star is the star of the global triangulation (the triangulation of points in general position
                is unique), so every edge at an owned point is a global edge; the rank that owns the endpoint with the
                lower global index emits it.  The sorted union over the ranks IS the canonical edge list of a single-GPU
                run, byte for byte (tests/test_slab.py, tests/test_gpu_slab.py).
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _capi


def _backend_device(compute_device):
    """tensors handed to torch.distributed live on the GPU under NCCL, on the CPU under gloo"""
    if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl":
        return compute_device
    return torch.device("cpu")


PEER_CAP = 256     # uncertified simplices a rank may hand to its peers for the exact in-sphere certificate


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def _allreduce(t, op, comm_dev):
    w, _ = _world()
    if w == 1:
        return t
    x = t.to(comm_dev)
    dist.all_reduce(x, op=op)
    return x.to(t.device)


def _allgather_rows(t, comm_dev):
    """variable-length all-gather of a [n, k] tensor: list of world tensors (on t.device)"""
    w, _ = _world()
    if w == 1:
        return [t]
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=comm_dev)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=comm_dev)
    pad[:t.shape[0]] = t.to(comm_dev)
    out = [torch.zeros_like(pad) for _ in range(w)]
    dist.all_gather(out, pad)
    return [o[:s].to(t.device) for o, s in zip(out, sizes)]


def _exchange(send, comm_dev):
    """send[r] = tensor for rank r (None / empty allowed); returns the list of tensors received from every rank.
    Peer-to-peer isend / irecv pairs (NCCL: over NVLink); sizes first."""
    w, me = _world()
    ref = next(t for t in send if t is not None)
    cnt = torch.tensor([0 if (t is None or r == me) else t.shape[0] for r, t in enumerate(send)], dtype=torch.int64, device=comm_dev)
    allc = [torch.zeros_like(cnt) for _ in range(w)]
    dist.all_gather(allc, cnt)
    counts = torch.stack(allc).cpu().numpy()        # counts[q, r] = rows q sends to r
    ops, bufs = [], [None] * w
    keep = []
    for r in range(w):
        if r == me:
            continue
        if counts[me, r] > 0:
            s = send[r].to(comm_dev).contiguous()
            keep.append(s)
            ops.append(dist.P2POp(dist.isend, s, r))
        if counts[r, me] > 0:
            bufs[r] = torch.empty((int(counts[r, me]),) + tuple(ref.shape[1:]), dtype=ref.dtype, device=comm_dev)
            ops.append(dist.P2POp(dist.irecv, bufs[r], r))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return [None if b is None else b.to(ref.device) for b in bufs], int(counts[me].sum()), int(counts[:, me].sum())


def _hash64(g):
    """splitmix-style hash of int64 global indices (torch int64 arithmetic wraps like uint64)"""
    def lsr(x, k):
        return (x >> k) & ((1 << (64 - k)) - 1)
    z = g * (-7046029254386353131) + 0x632BE59BD9B4E019
    z = (z ^ lsr(z, 30)) * (-4658895280553007687)
    z = (z ^ lsr(z, 27)) * (-7723592293110705685)
    return z ^ lsr(z, 31)


def partition_by_axis(points, world, rank, axis=0):
    """helper for callers that hold the WHOLE set: the rows of `points` (torch [n, dim]) that fall into slab `rank` of `world`
    equal-count slabs along `axis`, with their global indices.  Cuts are midpoints between neighbours in sorted order, so no
    point sits on a cut."""
    x = points[:, axis]
    order = torch.argsort(x)
    n = x.shape[0]
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    idx = torch.sort(order[lo:hi]).values
    return points[idx].contiguous(), idx.to(torch.int64)


class SlabResult:
    def __init__(self, edges, info):
        self.edges = edges      # this rank's part of the canonical list: uint32 [m, 2] global input indices, sorted
        self.info = info


def gather_edges(part):
    """concatenate the ranks' parts (disjoint by construction) and sort: the canonical edge list on every rank"""
    w, _ = _world()
    if w == 1:
        return part
    obj = [None] * w
    dist.all_gather_object(obj, part)
    e = np.concatenate(obj, axis=0)
    key = (e[:, 0].astype(np.uint64) << np.uint64(32)) | e[:, 1].astype(np.uint64)
    return e[np.argsort(key, kind="stable")]


def delaunay_slab(lib, points, global_index, device=0, axis=0, coarse_div=16, halo_spacings=4.0, max_rounds=24, verbose=False):
    """This rank's slab of one global triangulation.  points: torch float64 [n, dim] on the compute device (cuda:device for the
    product library; cpu for the kernel emulation in tests), all within this rank's range along `axis` (ranges of different
    ranks must not overlap); global_index: int64 [n], the input index of every point in the global set."""
    world, me = _world()
    verbose = verbose or bool(os.environ.get("VOR_SLAB_VERBOSE"))
    dim = int(points.shape[1])
    cdev = _backend_device(points.device)
    n_own = int(points.shape[0])
    pts = points.contiguous()
    gidx = global_index.to(points.device).to(torch.int64).contiguous()
    info = {"world": world, "rank": me, "own_points": n_own}

    def chk(st):
        if st not in (0, 3):
            raise _capi.VorError(st, lib.vor_last_error().decode())

    # ---- 1. global bounds and the 1.5x rule of bounding_sphere (geometry.rs:132-140)
    lo, hi = np.full(dim, np.inf), np.full(dim, -np.inf)
    if n_own:
        chk(lib.vor_slab_local_bounds(dim, C.c_void_p(pts.data_ptr()), n_own, device, lo.ctypes.data_as(_capi.dp), hi.ctypes.data_as(_capi.dp)))
    own_lo, own_hi = float(lo[axis]), float(hi[axis])
    glo = _allreduce(torch.from_numpy(lo.copy()), dist.ReduceOp.MIN, cdev).numpy().copy()
    ghi = _allreduce(torch.from_numpy(hi.copy()), dist.ReduceOp.MAX, cdev).numpy().copy()
    out = C.c_uint64(0)
    if n_own:
        chk(lib.vor_slab_count_outside(dim, C.c_void_p(pts.data_ptr()), n_own, device, glo.ctypes.data_as(_capi.dp), ghi.ctypes.data_as(_capi.dp), C.byref(out)))
    outside = int(_allreduce(torch.tensor([out.value], dtype=torch.int64), dist.ReduceOp.SUM, cdev).item())
    n_global = int(_allreduce(torch.tensor([n_own], dtype=torch.int64), dist.ReduceOp.SUM, cdev).item())

    # ---- ranges: cuts halfway between neighbouring ranks' extreme points; the outer ranks reach to the ends of the box
    ext = torch.tensor([[own_lo, own_hi]], dtype=torch.float64)
    exts = torch.cat(_allgather_rows(ext, cdev)).numpy()
    nonempty = [r for r in range(world) if np.isfinite(exts[r, 0])]
    for a, b in zip(nonempty[:-1], nonempty[1:]):
        if not exts[a, 1] < exts[b, 0]:
            raise ValueError("slab mode: the ranks' points must be partitioned along the axis (rank order = coordinate order)")
    my_lo, my_hi = float(glo[axis]), float(ghi[axis])
    if n_own:
        k = nonempty.index(me)
        if k > 0:
            my_lo = 0.5 * (exts[nonempty[k - 1], 1] + own_lo)
        if k + 1 < len(nonempty):
            my_hi = 0.5 * (own_hi + exts[nonempty[k + 1], 0])

    # ---- 2. coarse sample, identical on every rank
    is_coarse = (_hash64(gidx) & 0x7fffffff) % coarse_div == 0 if world > 1 else torch.zeros(n_own, dtype=torch.bool, device=pts.device)
    cparts = _allgather_rows(torch.cat([pts[is_coarse], gidx[is_coarse].to(torch.float64).view(-1, 1)], dim=1), cdev)
    cfrom = torch.cat([torch.full((p.shape[0],), r, dtype=torch.int64) for r, p in enumerate(cparts)])
    call = torch.cat(cparts)
    c_pts, c_g = call[:, :dim].contiguous(), call[:, dim].to(torch.int64)
    fine_pts, fine_g = pts[~is_coarse].contiguous(), gidx[~is_coarse]
    fine_x = fine_pts[:, axis]

    # ---- tree on the global bounds
    h = _capi.tree_p()
    hint = int(c_pts.shape[0] + 1.3 * fine_pts.shape[0] + 1024)
    chk(lib.vor_tree_create_bounds(dim, glo.ctypes.data_as(_capi.dp), ghi.ctypes.data_as(_capi.dp), outside, hint, device, None, C.byref(h)))
    gmap, owned = [], []                      # per local input index: global index, owned flag

    def insert(p, g, own_flags):
        if p.shape[0] == 0:
            return
        p = p.contiguous()
        if p.is_cuda:
            # the tree works on its own stream: what torch (or NCCL, through torch's stream) is still writing must have landed
            torch.cuda.current_stream(p.device).synchronize()
        chk(lib.vor_tree_insert_device(h, C.c_void_p(p.data_ptr()), int(p.shape[0]), 1))
        gmap.append(g.cpu().numpy().astype(np.int64))
        owned.append(own_flags)

    try:
        insert(c_pts, c_g, (cfrom == me).numpy().astype(np.uint8))
        insert(fine_pts, fine_g, np.ones(fine_pts.shape[0], dtype=np.uint8))

        # ---- 3./4. halo rounds: request region -> peers send what they have not sent yet -> insert -> certify.
        # Region of a rank = { axis coordinate in [xlo, xhi] }  u  { within `shell` of a lateral face of the data box }.
        # The shell exists for the hull: the circumsphere of a simplex on the hull is nearly a plane, its (empty) cap inside
        # the data box is thin but WIDE -- (2r/n)^(1/4) of the box for a sphere of radius r -- so hull simplices would ask
        # for a range far beyond their slab; all they miss are the few points right under the lateral faces.
        vol = float(np.prod(np.maximum(ghi - glo, 1e-300)))
        spacing = (vol / max(n_global, 1)) ** (1.0 / dim)
        lat = [k for k in range(dim) if k != axis]
        if fine_pts.shape[0]:
            latdist = torch.stack([torch.minimum(fine_pts[:, k] - float(glo[k]), float(ghi[k]) - fine_pts[:, k]) for k in lat]).min(dim=0).values
        else:
            latdist = torch.zeros(0, dtype=torch.float64, device=pts.device)
        shell_max = 16.0 * spacing
        want = [my_lo - halo_spacings * spacing, my_hi + halo_spacings * spacing, 2.0 * spacing] if world > 1 else [my_lo, my_hi, 0.0]
        have = [my_lo, my_hi, 0.0]            # region in which this tree holds every global point
        sent = {}                             # per peer: mask of my fine points already sent to it
        rounds = sent_rows = recv_rows = peer_certified = 0
        while True:
            want[0], want[1] = max(want[0], float(glo[axis])), min(want[1], float(ghi[axis]))
            if world > 1:
                req = torch.cat(_allgather_rows(torch.tensor([want], dtype=torch.float64), cdev)).numpy()
                send = [None] * world
                for r in range(world):
                    if r == me:
                        continue
                    done = sent.get(r)
                    if done is None:
                        done = torch.zeros(fine_pts.shape[0], dtype=torch.bool, device=pts.device)
                    m = (((fine_x >= req[r, 0]) & (fine_x <= req[r, 1])) | (latdist <= req[r, 2])) & ~done
                    send[r] = torch.cat([fine_pts[m], fine_g[m].to(torch.float64).view(-1, 1)], dim=1)
                    sent[r] = done | m
                got, ns, nr = _exchange(send, cdev)
                sent_rows += ns
                recv_rows += nr
                for g in got:
                    if g is not None and g.shape[0]:
                        insert(g[:, :dim].contiguous(), g[:, dim].to(torch.int64), np.zeros(g.shape[0], dtype=np.uint8))
            have = list(want)
            rounds += 1
            own_np = np.concatenate(owned) if owned else np.zeros(0, dtype=np.uint8)
            nunc, need = C.c_uint64(0), np.zeros(2)
            uverts, ureach = np.zeros((PEER_CAP, dim + 1, dim)), np.zeros((PEER_CAP, 2))
            if len(own_np):
                chk(lib.vor_tree_uncertified_slab(h, own_np.ctypes.data_as(C.POINTER(C.c_uint8)), len(own_np), axis, have[0], have[1], have[2],
                                                  uverts.ctypes.data_as(_capi.dp), ureach.ctypes.data_as(_capi.dp), PEER_CAP,
                                                  C.byref(nunc), need.ctypes.data_as(_capi.dp)))
            else:
                need[:] = have[:2]
            n_unc = int(nunc.value)
            # ---- the certificate that is not a ball.  The few simplices a rank cannot certify from its own ball (slivers on the
            # hull: a circumsphere of 1e3..1e5 box widths that f64 bounds loosely or not at all) go to every rank; each counts how many
            # of ITS fine points lie strictly inside (exact predicate).  None anywhere: the simplex is a simplex of the global
            # triangulation whatever its ball says.  Only when every rank's list fits (else the ranges are widened as before).
            if world > 1:
                fits = int(_allreduce(torch.tensor([1 if n_unc <= PEER_CAP else 0], dtype=torch.int64), dist.ReduceOp.MIN, cdev).item())
                total = int(_allreduce(torch.tensor([n_unc], dtype=torch.int64), dist.ReduceOp.SUM, cdev).item())
                if fits and total > 0:
                    mine_rows = torch.from_numpy(uverts[:n_unc].reshape(n_unc, -1).copy()) if n_unc else torch.zeros((0, (dim + 1) * dim), dtype=torch.float64)
                    lists = _allgather_rows(mine_rows, cdev)
                    counts = [int(x.shape[0]) for x in lists]
                    allv = torch.cat(lists).numpy() if total else np.zeros((0, (dim + 1) * dim))
                    inside = np.zeros(max(total, 1), dtype=np.uint64)
                    if total and fine_pts.shape[0]:
                        fp = fine_pts.contiguous()
                        if fp.is_cuda:
                            torch.cuda.current_stream(fp.device).synchronize()
                        allv = np.ascontiguousarray(allv, dtype=np.float64)
                        chk(lib.vor_points_in_spheres(dim, C.c_void_p(fp.data_ptr()), int(fp.shape[0]), allv.ctypes.data_as(_capi.dp), total, device,
                                                      inside.ctypes.data_as(_capi.u64p)))
                    tot_in = _allreduce(torch.from_numpy(inside[:max(total, 1)].astype(np.int64)), dist.ReduceOp.SUM, cdev).numpy()
                    off = int(sum(counts[:me]))
                    bad = [j for j in range(n_unc) if tot_in[off + j] > 0]
                    peer_certified += n_unc - len(bad)
                    n_unc = len(bad)
                    need[:] = have[:2]
                    for j in bad:
                        need[0], need[1] = min(need[0], ureach[j, 0]), max(need[1], ureach[j, 1])
            worst = int(_allreduce(torch.tensor([n_unc], dtype=torch.int64), dist.ReduceOp.MAX, cdev).item())
            if verbose:
                print(f"[slab {me}] round {rounds}: holds {sum(len(g) for g in gmap)} points, region {have}, uncertified {nunc.value} "
                      f"({n_unc} after asking the peers), need {need}", flush=True)
            if worst == 0:
                break
            if rounds >= max_rounds:
                raise RuntimeError("slab mode: certification did not converge")
            if n_unc:
                if have[2] < shell_max:
                    # first thicken the shell (cheap: a few percent of the points), and follow the need along the axis only a
                    # few spacings at a time
                    step = 3.0 * spacing
                    want = [max(float(need[0]), have[0] - step), min(float(need[1]), have[1] + step), 2.0 * have[2]]
                else:
                    grow = spacing
                    want = [min(have[0], float(need[0]) - grow) if need[0] < have[0] else have[0],
                            max(have[1], float(need[1]) + grow) if need[1] > have[1] else have[1], have[2]]
        info.update({"rounds": rounds, "peer_certified": peer_certified, "halo_rows_sent": sent_rows, "halo_rows_received": recv_rows, "coarse_points": int(c_pts.shape[0]),
                     "tree_points": int(sum(len(g) for g in gmap)), "region": have, "own_range": [my_lo, my_hi]})

        # ---- 5. edges at owned points, emitted by the owner of the endpoint with the lower global index (mapped, filtered and
        # sorted on the device: vor_tree_edges_slab)
        gm = np.ascontiguousarray(np.concatenate(gmap) if gmap else np.zeros(0, dtype=np.int64))
        own_np = np.ascontiguousarray(np.concatenate(owned) if owned else np.zeros(0, dtype=np.uint8))
        n_e = C.c_size_t()
        ptr = C.c_void_p()
        chk(lib.vor_tree_edges_slab(h, gm.ctypes.data_as(_capi.i64p), own_np.ctypes.data_as(C.POINTER(C.c_uint8)), len(gm), C.byref(ptr), C.byref(n_e)))
        out_e = np.asarray(_capi._HostBlock(lib, ptr.value, (n_e.value, 2), "<u4")) if n_e.value else np.zeros((0, 2), dtype=np.uint32)
    finally:
        lib.vor_tree_destroy(h)
    return SlabResult(out_e, info)

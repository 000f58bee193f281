"""ctypes declarations for include/voronoids_b200.h (shared by the product loader and the emulation tests)."""
import ctypes as C

import numpy as np

dp = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
szp = C.POINTER(C.c_size_t)
tree_p = C.c_void_p

STATUS = {0: "VOR_OK", 1: "VOR_ERR_NO_CONFLICT", 2: "VOR_ERR_DEGENERATE", 3: "VOR_ERR_DUPLICATE_POINT", 4: "VOR_ERR_CUDA",
          5: "VOR_ERR_OOM", 6: "VOR_ERR_CAPACITY", 7: "VOR_ERR_RANGE", 8: "VOR_ERR_OUTSIDE", 9: "VOR_ERR_INTERNAL", 10: "VOR_ERR_ARG"}

# every symbol include/voronoids_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "vor_tree_create": (C.c_int, [C.c_int, dp, C.c_size_t, C.c_int, C.POINTER(tree_p)]),
    "vor_tree_create_device": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(tree_p)]),
    "vor_tree_create_batch": (C.c_int, [C.c_int, dp, i64p, C.c_size_t, C.c_int, C.POINTER(tree_p)]),
    "vor_delaunay_batch": (C.c_int, [C.c_int, dp, i64p, C.c_size_t, C.POINTER(C.c_int), C.c_size_t, C.POINTER(tree_p), i64p]),
    "vor_tree_create_batch_device": (C.c_int, [C.c_int, C.c_void_p, i64p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(tree_p)]),
    "vor_tree_destroy": (None, [tree_p]),
    "vor_tree_insert": (C.c_int, [tree_p, dp, C.c_size_t, C.c_int]),
    "vor_tree_insert_device": (C.c_int, [tree_p, C.c_void_p, C.c_size_t, C.c_int]),
    "vor_tree_insert_batch": (C.c_int, [tree_p, dp, i64p]),
    "vor_tree_insert_batch_device": (C.c_int, [tree_p, C.c_void_p, i64p]),
    "vor_delaunay": (C.c_int, [C.c_int, dp, C.c_size_t, C.c_int, C.POINTER(tree_p)]),
    "vor_tree_counts": (C.c_int, [tree_p, u64p, u64p, u64p]),
    "vor_tree_edges": (C.c_int, [tree_p, u32p, C.c_size_t, szp]),
    "vor_tree_edges_host": (C.c_int, [tree_p, C.POINTER(C.c_void_p), szp]),
    "vor_host_free": (C.c_int, [C.c_void_p]),
    "vor_tree_edges_device": (C.c_int, [tree_p, C.POINTER(C.c_void_p), szp, u64p]),
    "vor_tree_export_simplices": (C.c_int, [tree_p, i32p, i32p, dp, dp, C.c_size_t, szp]),
    "vor_tree_locate": (C.c_int, [tree_p, dp, C.c_size_t, i32p, C.c_size_t, i32p]),
    "vor_tree_export_vertices": (C.c_int, [tree_p, dp, i64p, i32p, C.c_size_t, szp, szp]),
    "vor_make_queue": (C.c_int, [tree_p, dp, C.c_size_t, i64p, i32p, C.c_size_t, szp]),
    "vor_find_placement": (C.c_int, [i64p, i32p, C.c_size_t, u64p, C.c_int]),
    "vor_tree_check_delaunay": (C.c_int, [tree_p, C.POINTER(C.c_int), i32p]),
    "vor_debug_corrupt": (C.c_int, [tree_p, C.c_int]),
    "vor_tree_edges_slab": (C.c_int, [tree_p, i64p, C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "vor_slab_local_bounds": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_int, dp, dp]),
    "vor_slab_count_outside": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_int, dp, dp, u64p]),
    "vor_tree_create_bounds": (C.c_int, [C.c_int, dp, dp, C.c_uint64, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(tree_p)]),
    "vor_tree_certify_slab": (C.c_int, [tree_p, C.POINTER(C.c_uint8), C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_double, u64p, dp]),
    "vor_tree_uncertified_slab": (C.c_int, [tree_p, C.POINTER(C.c_uint8), C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_double, dp, dp,
                                            C.c_size_t, u64p, dp]),
    "vor_points_in_spheres": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, dp, C.c_size_t, C.c_int, u64p]),
    "vor_delaunay_batch_stream": (C.c_int, [C.c_int, C.c_void_p, C.c_int, i64p, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, u64p, u64p, C.c_void_p,
                                            C.c_void_p]),
    "vor_tree_super_simplex": (C.c_int, [tree_p, C.c_size_t, dp, dp, dp]),
    "vor_tree_stats": (C.c_int, [tree_p, u64p]),
    "vor_tree_profile": (C.c_int, [tree_p, dp]),
    "vor_circumsphere": (C.c_int, [C.c_int, dp, C.c_size_t, dp, dp, C.c_int]),
    "vor_in_sphere": (C.c_int, [C.c_int, dp, dp, dp, C.c_size_t, i32p, C.c_int]),
    "vor_bounding_sphere": (C.c_int, [C.c_int, dp, C.c_size_t, dp, dp, C.c_int]),
    "vor_predicates": (C.c_int, [C.c_int, dp, C.c_size_t, i32p, u64p, C.c_int]),
    "vor_sphere_filter": (C.c_int, [C.c_int, dp, C.c_double, dp, C.c_size_t, i32p, C.POINTER(C.c_float), C.c_int]),
    "vor_last_error": (C.c_char_p, []),
    "vor_kernel_launches": (C.c_uint64, []),
    "vor_release_memory": (None, []),
    "vor_set_option": (C.c_int, [C.c_char_p, C.c_double]),
    "vor_tree_set_stream": (None, [tree_p, C.c_void_p]),
}
N_STATS = 19
STAT_NAMES = ("rounds", "attempts", "winners", "owner_resets", "compactions", "stages", "walk_steps", "tests", "killed", "created",
              "exact_calls", "exact_zero", "duplicates", "simplex_slots", "aborted", "tests_completed", "sphere_undecided", "flagged", "slots")


CHUNK_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int64, C.POINTER(C.c_uint32), C.c_size_t)


def delaunay_batch_stream(lib, points, set_offsets, device=0, chunk_sets=0, chunk_points=0, on_chunk=None, dim=None):
    """vor_delaunay_batch_stream.  points: host float64 [n, dim] array, or an int device pointer (then pass dim).
    Returns (n_edges uint64 [n_sets], checksum64 uint64 [n_sets]); on_chunk(first_set, n_sets, first_point, edges[m,2])."""
    off = np.ascontiguousarray(set_offsets, dtype=np.int64)
    ns = len(off) - 1
    ne = np.zeros(ns, dtype=np.uint64)
    ck = np.zeros(ns, dtype=np.uint64)
    if isinstance(points, int):
        ptr, on_dev = C.c_void_p(points), 1
    else:
        p = np.ascontiguousarray(points, dtype=np.float64)
        dim = p.shape[1]
        ptr, on_dev = C.c_void_p(p.ctypes.data), 0
    cb = None
    if on_chunk is not None:
        def _cb(user, first_set, n_sets, first_point, edges, m):
            a = np.ctypeslib.as_array(edges, shape=(m, 2)).copy() if m else np.zeros((0, 2), dtype=np.uint32)
            on_chunk(int(first_set), int(n_sets), int(first_point), a)
        cb = CHUNK_CB(_cb)
    st = lib.vor_delaunay_batch_stream(dim, ptr, on_dev, off.ctypes.data_as(i64p), ns, device, chunk_sets, chunk_points,
                                       ne.ctypes.data_as(u64p), ck.ctypes.data_as(u64p), C.cast(cb, C.c_void_p) if cb else None, None)
    if st not in (0, 3):
        raise VorError(st, lib.vor_last_error().decode())
    return ne, ck


def delaunay_batch_devices(lib, points, set_offsets, devices):
    """vor_delaunay_batch: the sets in contiguous blocks over `devices` (one host thread each).  Returns (trees, shard):
    trees[d] is a Tree holding sets [shard[d], shard[d+1]) or None when its block is empty."""
    p, pp = as_f64(points)
    off = np.ascontiguousarray(set_offsets, dtype=np.int64)
    dev = (C.c_int * len(devices))(*devices)
    hs = (tree_p * len(devices))()
    shard = np.zeros(len(devices) + 1, dtype=np.int64)
    st = lib.vor_delaunay_batch(p.shape[1], pp, off.ctypes.data_as(i64p), len(off) - 1, dev, len(devices), hs, shard.ctypes.data_as(i64p))
    if st != 0:
        raise VorError(st, lib.vor_last_error().decode())
    trees = []
    for d in range(len(devices)):
        if not hs[d]:
            trees.append(None)
            continue
        t = Tree.__new__(Tree)
        t._lib, t._h, t.duplicates, t.dim = lib, tree_p(hs[d]), False, p.shape[1]
        t.n = int(off[shard[d + 1]] - off[shard[d]])
        t._off = off[shard[d]:shard[d + 1] + 1] - off[shard[d]]
        trees.append(t)
    return trees, shard


def find_placement(lib, offsets, ids, device=0):
    """scheduler::find_placement on a CSR queue: uint64 rounds (1-based)."""
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    idv = np.ascontiguousarray(ids, dtype=np.int32)
    n = off.size - 1
    out = np.zeros(max(n, 1), dtype=np.uint64)
    st = lib.vor_find_placement(off.ctypes.data_as(i64p), idv.ctypes.data_as(i32p) if idv.size else None, n, out.ctypes.data_as(u64p), device)
    if st != 0:
        raise VorError(st, lib.vor_last_error().decode())
    return out[:n]


class _HostBlock:
    """Owner of one result block of the library's host pool; numpy arrays made from it keep it alive through their .base."""

    def __init__(self, lib, ptr, shape, typestr):
        self._lib, self._ptr = lib, ptr
        self.__array_interface__ = {"data": (ptr, False), "shape": tuple(shape), "typestr": typestr, "version": 3}

    def __del__(self):
        try:
            self._lib.vor_host_free(C.c_void_p(self._ptr))
        except Exception:
            pass


def bind(lib):
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def as_f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


class VorError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"{STATUS.get(status, status)}: {text}")
        self.status = status


class Tree:
    """Thin RAII wrapper over a vor_tree handle (host-buffer entry points)."""

    def __init__(self, lib, points=None, device=0, dim=None, set_offsets=None, insert=True, one_shot=False):
        self._lib = lib
        self._h = tree_p()
        self.duplicates = False
        p, pp = as_f64(points)
        self.dim = p.shape[1] if dim is None else dim
        self.n = p.shape[0]
        if one_shot:
            # vor_delaunay: one host->device copy, DelaunayTree::new + insertion of every point (lib.rs:104-125)
            self._check(lib.vor_delaunay(self.dim, pp, self.n, device, C.byref(self._h)))
        elif set_offsets is None:
            self._check(lib.vor_tree_create(self.dim, pp, self.n, device, C.byref(self._h)))
            if insert:
                self.insert(p)
        else:
            off = np.ascontiguousarray(set_offsets, dtype=np.int64)
            self._off = off
            self._check(lib.vor_tree_create_batch(self.dim, pp, off.ctypes.data_as(i64p), len(off) - 1, device, C.byref(self._h)))
            if insert:
                self._check(lib.vor_tree_insert_batch(self._h, pp, off.ctypes.data_as(i64p)))

    @classmethod
    def from_device(cls, lib, ptr, n, dim, device=0, stream=None):
        """DelaunayTree::new + insertion of every point from a DEVICE buffer of n x dim float64 (row-major, contiguous):
        vor_tree_create_device + vor_tree_insert_device, no host copy of the coordinates."""
        self = cls.__new__(cls)
        self._lib = lib
        self._h = tree_p()
        self.duplicates = False
        self.dim, self.n = dim, n
        sp = C.c_void_p(stream) if stream else None
        self._check(lib.vor_tree_create_device(dim, C.c_void_p(ptr), n, device, sp, C.byref(self._h)))
        self._check(lib.vor_tree_insert_device(self._h, C.c_void_p(ptr), n, 1))
        return self

    def _check(self, st):
        if st == 3:
            self.duplicates = True
            return
        if st != 0:
            raise VorError(st, self._lib.vor_last_error().decode())

    def insert(self, points, mode=1):
        p, pp = as_f64(points)
        self._check(self._lib.vor_tree_insert(self._h, pp, p.shape[0], mode))

    def close(self):
        if self._h:
            self._lib.vor_tree_destroy(self._h)
            self._h = tree_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counts(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.vor_tree_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"vertices": a.value, "simplices": b.value, "max_simplex_id": c.value}

    def edges(self, pinned=True):
        """uint32 [m, 2].  pinned=True: zero-copy view of a block of the library's caching host allocator (recycled, hence
        resident; page-locked with VOR_PINNED_RESULTS=1); it goes back to the pool when the array is garbage collected.
        pinned=False: vor_tree_edges into a fresh numpy buffer."""
        n = C.c_size_t()
        if pinned and hasattr(self._lib, "vor_tree_edges_host"):
            ptr = C.c_void_p()
            self._check(self._lib.vor_tree_edges_host(self._h, C.byref(ptr), C.byref(n)))
            return np.asarray(_HostBlock(self._lib, ptr.value, (n.value, 2), "<u4"))
        self._check(self._lib.vor_tree_edges(self._h, None, 0, C.byref(n)))
        out = np.empty((n.value, 2), dtype=np.uint32)
        self._check(self._lib.vor_tree_edges(self._h, out.ctypes.data_as(u32p), n.value, C.byref(n)))
        return out

    def edge_checksum(self):
        n, ck, ptr = C.c_size_t(), C.c_uint64(), C.c_void_p()
        self._check(self._lib.vor_tree_edges_device(self._h, C.byref(ptr), C.byref(n), C.byref(ck)))
        return n.value, ck.value

    def simplices(self, circumspheres=False):
        n = C.c_size_t()
        self._check(self._lib.vor_tree_export_simplices(self._h, None, None, None, None, 0, C.byref(n)))
        m = self.dim + 1
        v = np.zeros((n.value, m), dtype=np.int32)
        nb = np.zeros((n.value, m), dtype=np.int32)
        c = np.zeros((n.value, self.dim)) if circumspheres else None
        r = np.zeros(n.value) if circumspheres else None
        self._check(self._lib.vor_tree_export_simplices(
            self._h, v.ctypes.data_as(i32p), nb.ctypes.data_as(i32p),
            c.ctypes.data_as(dp) if circumspheres else None, r.ctypes.data_as(dp) if circumspheres else None, n.value, C.byref(n)))
        return (v, nb, c, r) if circumspheres else (v, nb)

    def locate(self, points, cap=256):
        """Conflict regions of the query points: list of arrays of export indices (order of self.simplices())."""
        q, qp = as_f64(points)
        q = q.reshape(-1, self.dim)
        out = np.zeros((q.shape[0], cap), dtype=np.int32)
        cnt = np.zeros(q.shape[0], dtype=np.int32)
        self._check(self._lib.vor_tree_locate(self._h, qp, q.shape[0], out.ctypes.data_as(i32p), cap, cnt.ctypes.data_as(i32p)))
        if (cnt < 0).any():
            return self.locate(points, cap * 4)
        return [np.sort(out[i, :cnt[i]]) for i in range(q.shape[0])]

    def vertices(self):
        """(coords [n, dim], simp_off int64 [n + 1], simps int32): reference-ordered vertices and their incident simplices."""
        nv, ni = C.c_size_t(), C.c_size_t()
        self._check(self._lib.vor_tree_export_vertices(self._h, None, None, None, 0, C.byref(nv), C.byref(ni)))
        coords = np.zeros((nv.value, self.dim))
        off = np.zeros(nv.value + 1, dtype=np.int64)
        simps = np.zeros(max(ni.value, 1), dtype=np.int32)
        self._check(self._lib.vor_tree_export_vertices(self._h, coords.ctypes.data_as(dp), off.ctypes.data_as(i64p), simps.ctypes.data_as(i32p),
                                                       simps.size, C.byref(nv), C.byref(ni)))
        return coords, off, simps[:ni.value]

    def make_queue(self, points):
        """scheduler::make_queue: CSR (offsets int64 [n+1], ids int32) of the footprints, export indices."""
        q, qp = as_f64(points)
        q = q.reshape(-1, self.dim)
        n = q.shape[0]
        off = np.zeros(n + 1, dtype=np.int64)
        tot = C.c_size_t()
        self._check(self._lib.vor_make_queue(self._h, qp, n, off.ctypes.data_as(i64p), None, 0, C.byref(tot)))
        ids = np.zeros(max(tot.value, 1), dtype=np.int32)
        self._check(self._lib.vor_make_queue(self._h, qp, n, off.ctypes.data_as(i64p), ids.ctypes.data_as(i32p), ids.size, C.byref(tot)))
        return off, ids[:tot.value]

    def check_delaunay(self):
        ok = C.c_int()
        f = np.zeros(6, dtype=np.int32)
        self._check(self._lib.vor_tree_check_delaunay(self._h, C.byref(ok), f.ctypes.data_as(i32p)))
        return bool(ok.value), f

    def super_simplex(self, s=0):
        sv = np.zeros((self.dim + 1, self.dim))
        c = np.zeros(self.dim)
        r = C.c_double()
        self._check(self._lib.vor_tree_super_simplex(self._h, s, sv.ctypes.data_as(dp), c.ctypes.data_as(dp), C.byref(r)))
        return sv, c, r.value

    def stats(self):
        s = (C.c_uint64 * N_STATS)()
        self._check(self._lib.vor_tree_stats(self._h, s))
        return dict(zip(STAT_NAMES, [int(x) for x in s]))


def edge_checksum_host(edges):
    """Same order-independent checksum as vor_tree_edges_device, computed with numpy."""
    from .pointgen import splitmix64
    e = np.ascontiguousarray(edges, dtype=np.uint64)
    k = (e[:, 0] << np.uint64(32)) | e[:, 1]
    with np.errstate(over="ignore"):
        return int(np.sum(splitmix64(k), dtype=np.uint64))

"""Host-side mirror of the reference's public interface for the insertion path.

Rust rlib surface (/root/reference/src/delaunay_tree.rs)      ->  here
    DelaunayTree::<N,M>::new(points)              :390 / :545     DelaunayTree(points)            (also DelaunayTree.new)
    tree.add_points_to_tree(points)               :336            tree.add_points_to_tree(points)
    TreeUpdate::new + tree.insert_point(&update)  :710, :125      tree.insert_point(point)        (one-point round)
    tree.insert_points_parallel(&updates)         :213            tree.insert_points_parallel(points)
    tree.check_delaunay()                         :512 / :642     tree.check_delaunay()
    tree.max_simplex_id / vertices / simplices    :26-29          same names
PyO3 surface (/root/reference/src/lib.rs:12-134)
    voronoids.delaunay(points) -> PyDelauanyTree  :104-125        delaunay(points) -> PyDelauanyTree
    PyVertex.point/.simplex, PySimplex.vertices/.center/.radius/.neighbors   :12-60

Ids: vertices 0..dim = super simplex, dim+1..2dim+1 = ghost copies, 2(dim+1)+i = input point i (the reference's
sequential numbering, delaunay_tree.rs:173-174).  Simplex ids 1..dim+1 are the reference's ghost simplices
(:467-502); real simplices follow in engine order (the reference's own ids depend on its insertion order).
"""
from collections.abc import Mapping

import numpy as np

from . import _capi
from ._lib import lib


class PyVertex:
    """lib.rs:12-29"""
    __slots__ = ("point", "simplex")

    def __init__(self, point, simplex):
        self.point = point
        self.simplex = simplex

    def __repr__(self):
        return f"PyVertex(point={self.point}, simplex={self.simplex})"


class PySimplex:
    """lib.rs:31-60"""
    __slots__ = ("vertices", "center", "radius", "neighbors")

    def __init__(self, vertices, center, radius, neighbors):
        self.vertices = vertices
        self.center = center
        self.radius = radius
        self.neighbors = neighbors

    def __repr__(self):
        return f"PySimplex(vertices={self.vertices}, center={self.center}, radius={self.radius}, neighbors={self.neighbors})"


# ghost simplices of DelaunayTree::new: id -> vertex list (delaunay_tree.rs:467-502 / :606-632)
_GHOSTS = {3: {1: [4, 0, 1, 2], 2: [5, 0, 2, 3], 3: [6, 0, 3, 1], 4: [7, 1, 2, 3]},
           2: {1: [3, 0, 1], 2: [4, 0, 2], 3: [5, 1, 2]}}


# which ghost simplex lies behind a hull facet of the super simplex: the facet misses exactly one super vertex
_GHOST_OF_MISSING = {3: np.array([4, 2, 3, 1]), 2: np.array([3, 2, 1])}


class SimplexMap(Mapping):
    """`DelaunayTree.simplices` (lib.rs:87-101) as a lazy mapping id -> PySimplex over the exported numpy arrays: an item
    is built when it is asked for; nothing of size O(n) is materialised in Python objects (the PyO3 getter copies
    the whole DashMap into a dict on every access, README.md:24)."""

    def __init__(self, tree):
        self._tree = tree
        self._dim = tree.dim
        self._m = tree.dim + 1
        self._v, self._nb, self._c, self._r = tree.simplex_arrays()
        self._first = self._m + 1          # ids 1..M are the reference's ghost simplices
        self._n = len(self._v)
        if self._n == 1 and int(self._v.max()) < self._m:
            self._n = 0                    # nothing inserted yet: the only live simplex is the root, id 0 in the reference
        self._ghost_nb = None

    def _ghost_code(self):
        """[n, M] ghost id behind every hull facet (0 where the facet is interior), computed once with numpy."""
        if self._ghost_nb is None:
            hull = self._nb < 0
            tot = self._m * (self._m - 1) // 2       # 0 + 1 + ... + dim
            face_sum = self._v.sum(axis=1, keepdims=True) - self._v
            missing = np.where(hull, tot - face_sum, 0)
            self._ghost_nb = np.where(hull, _GHOST_OF_MISSING[self._dim][np.clip(missing, 0, self._m - 1)], 0)
        return self._ghost_nb

    def __len__(self):
        return self._n + self._m + (1 if self._n == 0 else 0)

    def __iter__(self):
        if self._n == 0:
            yield 0
        yield from range(1, self._m + 1)
        yield from range(self._first, self._first + self._n)

    def __contains__(self, k):
        try:
            k = int(k)
        except (TypeError, ValueError):
            return False
        return (1 <= k < self._first + self._n) or (k == 0 and self._n == 0)

    def __getitem__(self, k):
        k = int(k)
        ghosts = _GHOSTS[self._dim]
        if k == 0 and self._n == 0:   # nothing inserted yet: the root simplex is alive (delaunay_tree.rs:458-466)
            sv, cen, rad = self._tree._t.super_simplex()
            return PySimplex(list(range(self._m)), cen.tolist(), rad, list(ghosts))
        if 1 <= k <= self._m:
            if self._n == 0:
                return PySimplex(list(ghosts[k]), [0.0] * self._dim, 0.0, [0])
            rows = np.nonzero((self._ghost_code() == k).any(axis=1))[0]
            return PySimplex(list(ghosts[k]), [0.0] * self._dim, 0.0, (rows + self._first).tolist())
        i = k - self._first
        if not 0 <= i < self._n:
            raise KeyError(k)
        nb = self._nb[i]
        if (nb < 0).any():
            g = self._ghost_code()[i]
            neigh = [int(self._first + nb[q]) if nb[q] >= 0 else int(g[q]) for q in range(self._m)]
        else:
            neigh = (nb + self._first).tolist()
        return PySimplex(self._v[i].tolist(), self._c[i].tolist(), float(self._r[i]), neigh)


class VertexMap(Mapping):
    """`DelaunayTree.vertices` (lib.rs:73-85) as a lazy mapping id -> PyVertex over the exported CSR arrays."""

    def __init__(self, tree):
        self._dim = tree.dim
        self._m = tree.dim + 1
        self._coords, self._off, self._simps = tree._t.vertices()
        self._first = self._m + 1
        self._n = len(self._off) - 1
        self._ghost_inc = {}
        for gid, g in _GHOSTS[self._dim].items():
            for q in g:
                self._ghost_inc.setdefault(q, []).append(gid)

    def __len__(self):
        return self._n

    def __iter__(self):
        return iter(range(self._n))

    def __contains__(self, k):
        try:
            return 0 <= int(k) < self._n
        except (TypeError, ValueError):
            return False

    def __getitem__(self, k):
        k = int(k)
        if not 0 <= k < self._n:
            raise KeyError(k)
        inc = (self._simps[self._off[k]:self._off[k + 1]].astype(np.int64) + self._first).tolist()
        return PyVertex(self._coords[k].tolist(), inc + self._ghost_inc.get(k, []))


def _device_array(points):
    """(ptr, n, dim, device, keepalive) when `points` lives on a CUDA device (__cuda_array_interface__ or DLPack), else None."""
    cai = getattr(points, "__cuda_array_interface__", None)
    keep = points
    if cai is None and hasattr(points, "__dlpack__") and hasattr(points, "__dlpack_device__"):
        dev_type, _ = points.__dlpack_device__()
        if int(dev_type) != 2:                      # kDLCUDA
            return None
        import torch                                # plumbing only: DLPack capsule -> tensor -> pointer
        keep = torch.from_dlpack(points)
        cai = keep.__cuda_array_interface__
    if cai is None:
        return None
    shape, typestr, strides = tuple(cai["shape"]), cai["typestr"], cai.get("strides")
    if len(shape) != 2 or shape[1] not in (2, 3) or typestr not in ("<f8", "=f8", "|f8"):
        raise ValueError("device points must be float64 [n, 2] or [n, 3]")
    if strides is not None and tuple(strides) != (shape[1] * 8, 8):
        raise ValueError("device points must be C-contiguous")
    device = getattr(getattr(keep, "device", None), "index", None)
    if device is None:
        device = getattr(getattr(keep, "device", None), "id", 0) or 0
    return int(cai["data"][0]), int(shape[0]), int(shape[1]), int(device), keep


class DelaunayTree:
    """DelaunayTree<N,M> on the device (delaunay_tree.rs:24-30)."""

    def __init__(self, points, device=0, _tree=None, _dim=None):
        if _tree is not None and _dim is not None:
            self.dim = _dim
            self._t = _tree
        else:
            p = np.ascontiguousarray(points, dtype=np.float64)
            if p.ndim != 2 or p.shape[1] not in (2, 3):
                raise ValueError("points must be [n, 2] or [n, 3]")
            self.dim = p.shape[1]
            self._t = _tree if _tree is not None else _capi.Tree(lib(), p, device=device, insert=False)
        self._cache = None
        self._smap = self._vmap = None
        self._chunks = []   # inserted point arrays (kept by reference)

    new = classmethod(lambda cls, points, device=0: cls(points, device))

    # ---- mutation
    def add_points_to_tree(self, points):
        """delaunay_tree.rs:336-386"""
        self._cache = self._smap = self._vmap = None
        p = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, self.dim)
        self._t.insert(p, mode=1)
        self._chunks.append(p)

    insert_points_parallel = add_points_to_tree

    def insert_point(self, point):
        """TreeUpdate::new + insert_point (delaunay_tree.rs:710-739, :125-211) for one point."""
        self._cache = self._smap = self._vmap = None
        p = np.asarray(point, dtype=np.float64).reshape(1, self.dim)
        self._t.insert(p, mode=0)
        self._chunks.append(p)

    # ---- queries
    @property
    def duplicates_dropped(self):
        return self._t.duplicates

    @property
    def max_simplex_id(self):
        return self._t.counts()["max_simplex_id"]

    def counts(self):
        return self._t.counts()

    def locate(self, vertex):
        """DelaunayTree::locate (delaunay_tree.rs:33-58): sorted ids of the simplices in conflict with `vertex`
        (the ids used as keys of `.simplices`)."""
        first = self.dim + 2
        return [int(first + i) for i in self._t.locate(np.asarray(vertex, dtype=np.float64).reshape(1, self.dim))[0]]

    def voronoi(self):
        """Voronoi dual of the current triangulation: (vertices = circumcentres [n, N], ridges = pairs of adjacent
        simplex indices [m, 2]) over the live simplices in export order."""
        v, nb, c, r = self.simplex_arrays()
        i = np.repeat(np.arange(len(nb)), nb.shape[1])
        j = nb.reshape(-1)
        keep = (j >= 0) & (i < j)
        return c, np.stack([i[keep], j[keep]], axis=1)

    def check_delaunay(self):
        """delaunay_tree.rs:512-541 (local-Delaunay formulation, see include/voronoids_b200.h)"""
        return self._t.check_delaunay()[0]

    def edges(self):
        """Canonical Delaunay graph: sorted unique (lo, hi) input-index pairs, uint32 [m, 2]."""
        return self._t.edges()

    def stats(self):
        return self._t.stats()

    def super_simplex(self):
        return self._t.super_simplex()

    def simplex_arrays(self):
        """(vertices [n,M] reference ids, neighbors [n,M] export indices or -1, centers [n,N], radii [n])."""
        if self._cache is None:
            self._cache = self._t.simplices(circumspheres=True)
        return self._cache

    # ---- reference-shaped views: lazy mappings over numpy arrays (the PyO3 getters copy whole maps on every access)
    @property
    def simplices(self):
        """id -> PySimplex (lib.rs:87-101)"""
        if self._smap is None:
            self._smap = SimplexMap(self)
        return self._smap

    @property
    def vertices(self):
        """id -> PyVertex (lib.rs:73-85)"""
        if self._vmap is None:
            self._vmap = VertexMap(self)
        return self._vmap

    def close(self):
        self._t.close()


class PyDelauanyTree(DelaunayTree):
    """lib.rs:62-102 (the reference's spelling)."""


def delaunay(points, device=0):
    """voronoids.delaunay(points) (lib.rs:104-125): build the tree and insert every point.

    The reference inserts the first 1e5 points one by one and the rest through add_points_to_tree; both produce
    the (unique) Delaunay triangulation of points + super vertices, which is what the device rounds compute.
    """
    dev = _device_array(points)
    if dev is not None:
        # zero-copy input: points already on a CUDA device (__cuda_array_interface__ / DLPack): no host round trip
        ptr, n, dim, dev_index, keep = dev
        t = PyDelauanyTree(None, _tree=_capi.Tree.from_device(lib(), ptr, n, dim, device=dev_index), _dim=dim)
        t._device_points = keep
        return t
    p = np.ascontiguousarray(points, dtype=np.float64)
    if p.ndim != 2 or p.shape[1] not in (2, 3):
        raise ValueError("points must be [n, 2] or [n, 3]")
    t = PyDelauanyTree(p, device=device, _tree=_capi.Tree(lib(), p, device=device, one_shot=True))
    t._chunks.append(p)
    return t


class BatchResult:
    """Independent point sets triangulated in one device store (BASELINE.json config 5)."""

    def __init__(self, tree, offsets):
        self._t = tree
        self.offsets = np.asarray(offsets, dtype=np.int64)

    def edges(self, s=None):
        """All edges (global input indices) or those of set s (indices local to the set)."""
        e = self._t.edges()
        if s is None:
            return e
        lo, hi = self.offsets[s], self.offsets[s + 1]
        a = np.searchsorted(e[:, 0], lo, side="left")
        b = np.searchsorted(e[:, 0], hi, side="left")
        return (e[a:b] - np.uint32(lo)).astype(np.uint32)

    def check_delaunay(self):
        return self._t.check_delaunay()[0]

    def stats(self):
        return self._t.stats()

    def close(self):
        self._t.close()


def delaunay_batch(point_sets, device=0):
    """Triangulate a list of independent [n_s, dim] point sets together; returns a BatchResult."""
    sets = [np.ascontiguousarray(p, dtype=np.float64) for p in point_sets]
    off = np.zeros(len(sets) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(p) for p in sets])
    allp = np.concatenate(sets, axis=0)
    t = _capi.Tree(lib(), allp, device=device, set_offsets=off)
    return BatchResult(t, off)


def delaunay_batch_stream(points, set_offsets, device=0, chunk_sets=0, chunk_points=0, on_chunk=None):
    """Batches far beyond one device store (BASELINE.json configs[4]: 8,192 sets x 100k points; the batch loop of
    examples/parallel_insert.rs:56-78): the sets are triangulated chunk by chunk (vor_delaunay_batch_stream), the copy of
    the next chunk overlapping the rounds of the current one.  `points`: float64 [n, dim] on the host, or a device array
    (__cuda_array_interface__ / DLPack) holding all sets.  Returns (n_edges, checksum64) per set as uint64 arrays;
    on_chunk(first_set, n_sets, first_point, edges) receives every chunk's canonical edge list (chunk-local indices)."""
    dev = _device_array(points)
    if dev is not None:
        ptr, n, dim, dev_index, keep = dev
        return _capi.delaunay_batch_stream(lib(), ptr, set_offsets, device=dev_index, chunk_sets=chunk_sets, chunk_points=chunk_points,
                                           on_chunk=on_chunk, dim=dim)
    return _capi.delaunay_batch_stream(lib(), np.ascontiguousarray(points, dtype=np.float64), set_offsets, device=device, chunk_sets=chunk_sets,
                                       chunk_points=chunk_points, on_chunk=on_chunk)

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


class DelaunayTree:
    """DelaunayTree<N,M> on the device (delaunay_tree.rs:24-30)."""

    def __init__(self, points, device=0, _tree=None):
        p = np.ascontiguousarray(points, dtype=np.float64)
        if p.ndim != 2 or p.shape[1] not in (2, 3):
            raise ValueError("points must be [n, 2] or [n, 3]")
        self.dim = p.shape[1]
        self._t = _tree if _tree is not None else _capi.Tree(lib(), p, device=device, insert=False)
        self._cache = None
        self._chunks = []   # inserted point arrays (kept by reference; concatenated only if .vertices is asked for)

    new = classmethod(lambda cls, points, device=0: cls(points, device))

    # ---- mutation
    def add_points_to_tree(self, points):
        """delaunay_tree.rs:336-386"""
        self._cache = None
        p = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, self.dim)
        self._t.insert(p, mode=1)
        self._chunks.append(p)

    insert_points_parallel = add_points_to_tree

    def insert_point(self, point):
        """TreeUpdate::new + insert_point (delaunay_tree.rs:710-739, :125-211) for one point."""
        self._cache = None
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

    # ---- reference-shaped views (built on demand, like the PyO3 getters that copy the maps on every access)
    @property
    def simplices(self):
        m = self.dim + 1
        v, nb, c, r = self.simplex_arrays()
        first = m + 1  # ids 1..M are the ghosts
        ghosts = _GHOSTS[self.dim]
        face_to_ghost = {frozenset(g[1:]): gid for gid, g in ghosts.items()}
        out = {}
        ghost_nb = {gid: [] for gid in ghosts}
        vl, nl, cl, rl = v.tolist(), nb.tolist(), c.tolist(), r.tolist()
        for i in range(len(vl)):
            neigh = []
            for k in range(m):
                j = nl[i][k]
                if j >= 0:
                    neigh.append(first + j)
                else:
                    gid = face_to_ghost[frozenset(vl[i][q] for q in range(m) if q != k)]
                    neigh.append(gid)
                    ghost_nb[gid].append(first + i)
            out[first + i] = PySimplex(vl[i], cl[i], rl[i], neigh)
        if not vl:  # nothing inserted yet: the root simplex 0 is alive (delaunay_tree.rs:458-466)
            sv, cen, rad = self._t.super_simplex()
            out[0] = PySimplex(list(range(m)), cen.tolist(), rad, list(ghosts))
            ghost_nb = {gid: [0] for gid in ghosts}
        for gid, g in ghosts.items():
            out[gid] = PySimplex(list(g), [0.0] * self.dim, 0.0, ghost_nb[gid])
        return out

    @property
    def vertices(self):
        m = self.dim + 1
        sv = self._t.super_simplex()[0]
        n_real = self._t.counts()["vertices"] - 2 * m
        first = m + 1
        # Vertex.simplex from the device (vor_tree_export_vertices): CSR of export indices per reference vertex id
        _, off, simps = self._t.vertices()
        ids = (simps.astype(np.int64) + first).tolist()
        inc = {q: ids[off[q]:off[q + 1]] for q in range(len(off) - 1) if off[q + 1] > off[q]}
        ghosts = _GHOSTS[self.dim]
        for gid, g in ghosts.items():
            for q in g:
                inc.setdefault(q, []).append(gid)
        out = {}
        # super + ghost coordinates (delaunay_tree.rs:407-412 / :559-566)
        ghost_of = {3: [0, 0, 0, 1], 2: [0, 1, 2]}[self.dim]
        for k in range(m):
            out[k] = PyVertex(sv[k].tolist(), inc.get(k, []))
            out[m + k] = PyVertex(sv[ghost_of[k]].tolist(), inc.get(m + k, []))
        pts = self._points()
        for i in range(n_real):
            out[2 * m + i] = PyVertex(pts[i].tolist(), inc.get(2 * m + i, []))
        return out

    def _points(self):
        if len(self._chunks) != 1:
            self._chunks = [np.concatenate(self._chunks, axis=0) if self._chunks else np.zeros((0, self.dim))]
        return self._chunks[0]

    def close(self):
        self._t.close()


class PyDelauanyTree(DelaunayTree):
    """lib.rs:62-102 (the reference's spelling)."""


def delaunay(points, device=0):
    """voronoids.delaunay(points) (lib.rs:104-125): build the tree and insert every point.

    The reference inserts the first 1e5 points one by one and the rest through add_points_to_tree; both produce
    the (unique) Delaunay triangulation of points + super vertices, which is what the device rounds compute.
    """
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

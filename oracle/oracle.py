"""ctypes front end of the CPU oracle -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (the product path never does).

  exact_*      exact-predicate Bowyer-Watson + predicates  (oracle/exact_bw.c, predicates.c)
  ref_*        float restatement of kazewong/Voronoids      (oracle/refcpu.cpp, ref_geometry.h)
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("predicates.c", "exact_bw.c", "ref_geometry.c", "ref_geometry.h", "refcpu.cpp", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip, u32p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.vo_bw_create.restype = C.c_void_p
        L.vo_bw_create.argtypes = [C.c_int, dp, C.c_int, dp]
        L.vo_bw_destroy.argtypes = [C.c_void_p]
        L.vo_bw_error.argtypes = [C.c_void_p]
        L.vo_bw_stats.argtypes = [C.c_void_p, u64p]
        L.vo_bw_validate.argtypes = [C.c_void_p]
        L.vo_bw_simplices.restype = C.c_uint64
        L.vo_bw_simplices.argtypes = [C.c_void_p, ip, C.c_uint64]
        L.vo_bw_edges.restype = C.c_uint64
        L.vo_bw_edges.argtypes = [C.c_void_p, u32p, C.c_uint64]
        L.vo_pred_counters.argtypes = [u64p, u64p, u64p, C.c_int]
        for nm in ("vo_orient2d_batch", "vo_orient3d_batch", "vo_incircle_batch", "vo_insphere_batch"):
            getattr(L, nm).argtypes = [dp, C.c_int, ip, C.c_int]
        L.vo_ref_in_sphere.argtypes = [C.c_int, dp, dp, C.c_double]
        L.vo_ref_circumsphere.argtypes = [C.c_int, dp, dp, dp]
        L.vo_ref_bounding_sphere.argtypes = [C.c_int, dp, C.c_long, dp, dp]
        L.vo_ref_super_simplex.argtypes = [C.c_int, dp, C.c_long, dp, dp, dp]
        L.vo_ref_create.restype = C.c_void_p
        L.vo_ref_create.argtypes = [C.c_int, dp, C.c_long]
        L.vo_ref_destroy.argtypes = [C.c_void_p]
        L.vo_ref_insert_sequential.argtypes = [C.c_void_p, dp, C.c_long, C.c_long]
        L.vo_ref_add_points_to_tree.argtypes = [C.c_void_p, dp, C.c_long, C.c_long, C.c_int]
        L.vo_ref_delaunay.restype = C.c_void_p
        L.vo_ref_delaunay.argtypes = [C.c_int, dp, C.c_long, C.c_int, ip]
        L.vo_ref_counts.argtypes = [C.c_void_p, u64p]
        L.vo_ref_check_delaunay.argtypes = [C.c_void_p]
        L.vo_ref_edges.restype = C.c_uint64
        L.vo_ref_edges.argtypes = [C.c_void_p, u32p, C.c_uint64]
        L.vo_ref_placement.argtypes = [C.c_void_p, dp, C.c_long, u64p]
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


# ---------------------------------------------------------------- geometry (float restatement)
def ref_circumsphere(verts):
    v, vp = _d(verts)
    n = v.shape[1]
    c = np.zeros(n)
    r = C.c_double()
    rc = lib().vo_ref_circumsphere(n, vp, c.ctypes.data_as(C.POINTER(C.c_double)), C.byref(r))
    if rc:
        raise ArithmeticError("singular LU (reference: unwrap panic, geometry.rs:49)")
    return c, r.value


def ref_in_sphere(vertex, center, radius):
    v, vp = _d(vertex)
    c, cp = _d(center)
    return bool(lib().vo_ref_in_sphere(len(v), vp, cp, float(radius)))


def ref_bounding_sphere(pts):
    p, pp = _d(pts)
    n = p.shape[1]
    c = np.zeros(n)
    r = C.c_double()
    lib().vo_ref_bounding_sphere(n, pp, p.shape[0], c.ctypes.data_as(C.POINTER(C.c_double)), C.byref(r))
    return c, r.value


def ref_super_simplex(pts):
    """(super vertices [M,N], center, 10x radius) exactly as DelaunayTree::new computes them."""
    p, pp = _d(pts)
    n = p.shape[1]
    sup = np.zeros((n + 1, n))
    c = np.zeros(n)
    r = C.c_double()
    dp = C.POINTER(C.c_double)
    lib().vo_ref_super_simplex(n, pp, p.shape[0], sup.ctypes.data_as(dp), c.ctypes.data_as(dp), C.byref(r))
    return sup, c, r.value


# ---------------------------------------------------------------- predicates
def _batch(name, rows, width, exact_only):
    a, ap = _d(rows)
    a = a.reshape(-1, width)
    out = np.zeros(a.shape[0], dtype=np.int32)
    getattr(lib(), name)(ap, a.shape[0], out.ctypes.data_as(C.POINTER(C.c_int)), int(exact_only))
    return out


def orient2d(rows, exact_only=False):
    return _batch("vo_orient2d_batch", rows, 6, exact_only)


def orient3d(rows, exact_only=False):
    return _batch("vo_orient3d_batch", rows, 12, exact_only)


def incircle(rows, exact_only=False):
    return _batch("vo_incircle_batch", rows, 8, exact_only)


def insphere(rows, exact_only=False):
    return _batch("vo_insphere_batch", rows, 15, exact_only)


def pred_counters(reset=False):
    f, e, z = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib().vo_pred_counters(C.byref(f), C.byref(e), C.byref(z), int(reset))
    return {"filter": f.value, "exact": e.value, "zero": z.value}


# ---------------------------------------------------------------- canonical edges
def edge_sha256(edges):
    e = np.ascontiguousarray(edges, dtype="<u4")
    return hashlib.sha256(e.tobytes()).hexdigest()


class ExactDelaunay:
    """Exact Bowyer-Watson of P u S; vertex ids 0..M-1 super, M+i = input point i."""

    def __init__(self, pts, super_vertices=None):
        p, pp = _d(pts)
        self.dim = p.shape[1]
        self.n = p.shape[0]
        if super_vertices is None:
            super_vertices = ref_super_simplex(p)[0]
        s, sp = _d(super_vertices)
        pred_counters(reset=True)
        self._h = lib().vo_bw_create(self.dim, pp, self.n, sp)
        self.pred = pred_counters()
        err = lib().vo_bw_error(self._h)
        if err:
            raise RuntimeError({1: "duplicate point / empty conflict region", 2: "point outside the super simplex"}[err])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vo_bw_destroy(self._h)
            self._h = None

    def stats(self):
        out = (C.c_uint64 * 5)()
        lib().vo_bw_stats(self._h, out)
        return dict(zip(("live", "created", "killed", "walk_steps", "tests"), [int(x) for x in out]))

    def validate(self):
        return lib().vo_bw_validate(self._h)

    def simplices(self):
        m = lib().vo_bw_simplices(self._h, None, 0)
        out = np.zeros((m, self.dim + 1), dtype=np.int32)
        lib().vo_bw_simplices(self._h, out.ctypes.data_as(C.POINTER(C.c_int)), m)
        return out

    def edges(self):
        m = lib().vo_bw_edges(self._h, None, 0)
        out = np.zeros((m, 2), dtype=np.uint32)
        lib().vo_bw_edges(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), m)
        return out


class RefDelaunay:
    """Float restatement of the reference (oracle/refcpu.cpp). mode: 'delaunay' = lib.rs:104-125."""

    def __init__(self, pts, mode="delaunay", n_seq=None, nthreads=0):
        p, pp = _d(pts)
        self.dim = p.shape[1]
        self.n = p.shape[0]
        L = lib()
        err = C.c_int(0)
        if mode == "delaunay":
            self._h = L.vo_ref_delaunay(self.dim, pp, self.n, nthreads, C.byref(err))
            e = err.value
        elif mode == "new":  # DelaunayTree::new only
            self._h = L.vo_ref_create(self.dim, pp, self.n)
            e = 0
        else:
            self._h = L.vo_ref_create(self.dim, pp, self.n)
            n_seq = self.n if n_seq is None else n_seq
            e = L.vo_ref_insert_sequential(self._h, pp, n_seq, 0)
            if not e and n_seq < self.n:
                rest = np.ascontiguousarray(p[n_seq:])
                e = L.vo_ref_add_points_to_tree(self._h, rest.ctypes.data_as(C.POINTER(C.c_double)), self.n - n_seq, n_seq, nthreads)
        self.err = e

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vo_ref_destroy(self._h)
            self._h = None

    def counts(self):
        out = (C.c_uint64 * 4)()
        lib().vo_ref_counts(self._h, out)
        return dict(zip(("vertices", "live", "max_simplex_id", "rounds"), [int(x) for x in out]))

    def check_delaunay(self):
        return bool(lib().vo_ref_check_delaunay(self._h))

    def edges(self):
        m = lib().vo_ref_edges(self._h, None, 0)
        out = np.zeros((m, 2), dtype=np.uint32)
        lib().vo_ref_edges(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), m)
        return out

    def placement(self, pts):
        p, pp = _d(pts)
        out = np.zeros(p.shape[0], dtype=np.uint64)
        lib().vo_ref_placement(self._h, pp, p.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out

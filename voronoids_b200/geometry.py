"""Mirror of the reference's `pub mod geometry` (/root/reference/src/geometry.rs) on the device, batched.

    circumsphere(vertices)          geometry.rs:58-87   one simplex [M,N] or a batch [n,M,N]
    in_sphere(vertex, center, radius)   geometry.rs:91-97
    bounding_sphere(points)         geometry.rs:99-142
and the exact predicates the engine adds (no reference counterpart): orient2d/orient3d/incircle/insphere.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._lib import lib


def _check(st):
    if st != 0:
        raise _capi.VorError(st, lib().vor_last_error().decode())


def circumsphere(vertices, device=0):
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    single = v.ndim == 2
    if single:
        v = v[None]
    n, m, d = v.shape
    c = np.zeros((n, d))
    r = np.zeros(n)
    _check(lib().vor_circumsphere(d, v.ctypes.data_as(_capi.dp), n, c.ctypes.data_as(_capi.dp), r.ctypes.data_as(_capi.dp), device))
    return (c[0], float(r[0])) if single else (c, r)


def in_sphere(vertex, center, radius, device=0):
    p = np.ascontiguousarray(vertex, dtype=np.float64)
    single = p.ndim == 1
    p = np.atleast_2d(p)
    c = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(center, dtype=np.float64)), p.shape))
    r = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(np.asarray(radius, dtype=np.float64)), (p.shape[0],)))
    out = np.zeros(p.shape[0], dtype=np.int32)
    _check(lib().vor_in_sphere(p.shape[1], p.ctypes.data_as(_capi.dp), c.ctypes.data_as(_capi.dp), r.ctypes.data_as(_capi.dp), p.shape[0],
                               out.ctypes.data_as(_capi.i32p), device))
    return bool(out[0]) if single else out.astype(bool)


def bounding_sphere(points, device=0):
    p = np.ascontiguousarray(points, dtype=np.float64)
    c = np.zeros(p.shape[1])
    r = C.c_double()
    _check(lib().vor_bounding_sphere(p.shape[1], p.ctypes.data_as(_capi.dp), p.shape[0], c.ctypes.data_as(_capi.dp), C.byref(r), device))
    return c, r.value


_KIND = {"orient2d": (0, 6), "orient3d": (1, 12), "incircle": (2, 8), "insphere": (3, 15)}


def predicate(kind, rows, device=0, return_exact_count=False):
    k, w = _KIND[kind]
    a = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, w)
    out = np.zeros(a.shape[0], dtype=np.int32)
    ne = C.c_uint64()
    _check(lib().vor_predicates(k, a.ctypes.data_as(_capi.dp), a.shape[0], out.ctypes.data_as(_capi.i32p), C.byref(ne), device))
    return (out, ne.value) if return_exact_count else out

"""Shared material for the cached-circumsphere FILTER (voronoids_b200/csrc/sphere.cuh): adversarial simplex + query rows
and the check that every verdict the filter certifies (+1 / -1) agrees with the exact predicate.

The reference tests `dist^2 < r*r` on a cached float centre/radius (delaunay_tree.rs:11-16, geometry.rs:91-97) and has
no error control; here the cached sphere may only answer when it is provably right."""
import ctypes as C

import numpy as np

from voronoids_b200 import _capi


def sphere_filter(lib, dim, origin, reach, rows):
    a = np.ascontiguousarray(rows, dtype=np.float64)
    n = a.shape[0]
    out = np.zeros(n, dtype=np.int32)
    blocks = np.zeros((n, 5), dtype=np.float32)
    org = np.ascontiguousarray(origin, dtype=np.float64)
    st = lib.vor_sphere_filter(dim, org.ctypes.data_as(_capi.dp), float(reach), a.ctypes.data_as(_capi.dp), n, out.ctypes.data_as(_capi.i32p),
                               blocks.ctypes.data_as(C.POINTER(C.c_float)), 0)
    assert st == 0
    return out, blocks


def _circum(simp):
    """float64 circumcentre / radius of one simplex (rows of points), good enough to aim queries at the sphere."""
    a = simp[1:] - simp[0]
    rhs = 0.5 * (a * a).sum(axis=1)
    try:
        c = np.linalg.solve(a, rhs)
    except np.linalg.LinAlgError:
        return None, None
    return simp[0] + c, float(np.sqrt((c * c).sum()))


def make_rows(dim, n, seed, scale=1.0, offset=0.0):
    """n rows [simplex (dim+1 points), query]: well-shaped, sliver, needle and tiny simplices; queries far inside / far
    outside / within 1e-3 ... 1e-16 (relative) of the sphere / equal to a vertex / a vertex nudged by a few ulps."""
    rng = np.random.default_rng(seed)
    M = dim + 1
    rows = []
    while len(rows) < n:
        kind = rng.integers(0, 5)
        simp = rng.random((M, dim))
        if kind == 1:      # sliver: last vertex almost in the hyperplane of the others
            w = rng.random(M - 1); w /= w.sum()
            simp[-1] = (w[:, None] * simp[:-1]).sum(axis=0) + rng.normal(size=dim) * 10.0 ** rng.uniform(-14, -3)
        elif kind == 2:    # tiny simplex somewhere in the unit box
            simp = rng.random(dim) + (simp - 0.5) * 10.0 ** rng.uniform(-9, -2)
        elif kind == 3:    # needle: two vertices almost coincide
            simp[1] = simp[0] + rng.normal(size=dim) * 10.0 ** rng.uniform(-12, -4)
        elif kind == 4:    # huge circumsphere (nearly flat, like hull simplices against a super vertex)
            simp[-1] = simp[0] + (simp[1] - simp[0]) * rng.uniform(0.2, 0.8) + rng.normal(size=dim) * 10.0 ** rng.uniform(-10, -5)
        simp = simp * scale + offset
        c, r = _circum(simp)
        qk = rng.integers(0, 6)
        if c is None or not np.isfinite(r) or qk == 0:
            q = rng.random(dim) * scale + offset
        elif qk in (1, 2, 3):
            # aim at the sphere near the simplex itself (a huge sphere is only ever queried next to its simplex:
            # the engine's queries lie inside the bounding box of the point set)
            d = simp.mean(axis=0) - c + rng.normal(size=dim) * 0.3 * np.abs(simp - simp.mean(axis=0)).max()
            d /= np.sqrt((d * d).sum())
            eps = 10.0 ** rng.uniform(-16, -3) * rng.choice([-1.0, 1.0])
            q = c + d * r * (1.0 + eps)
        elif qk == 4:
            q = simp[rng.integers(0, M)].copy()
        else:
            q = simp[rng.integers(0, M)].copy()
            for k in range(dim):
                for _ in range(rng.integers(0, 4)):
                    q[k] = np.nextafter(q[k], rng.choice([-np.inf, np.inf]))
        if np.abs(q - (offset + 0.5 * scale)).max() > 1.5 * scale:
            continue
        rows.append(np.concatenate([simp.reshape(-1), q]))
    return np.array(rows)


def exact_inside(oracle, dim, rows):
    """+1 strictly inside the circumsphere, 0 on it, -1 outside (orientation-independent); nan-free rows only."""
    if dim == 3:
        s = oracle.insphere(rows).astype(np.int64) * oracle.orient3d(rows[:, :12]).astype(np.int64)
        flat = oracle.orient3d(rows[:, :12]) == 0
    else:
        s = oracle.incircle(rows).astype(np.int64) * oracle.orient2d(rows[:, :6]).astype(np.int64)
        flat = oracle.orient2d(rows[:, :6]) == 0
    return s, flat


def check_filter(lib, oracle, dim, n, seed, scale=1.0, offset=0.0):
    rows = make_rows(dim, n, seed, scale, offset)
    origin = np.full(dim, offset + 0.5 * scale)
    reach = float(np.abs(rows.reshape(n, dim + 2, dim) - origin).sum(axis=2).max())
    verdict, blocks = sphere_filter(lib, dim, origin, reach, rows)
    want, flat = exact_inside(oracle, dim, rows)
    # a flat simplex has no circumsphere: the filter must not answer
    assert np.all(verdict[flat] == 0)
    ok = ~flat
    wrong_in = ok & (verdict > 0) & (want <= 0)
    wrong_out = ok & (verdict < 0) & (want > 0)
    assert not wrong_in.any(), f"filter certified INSIDE wrongly on rows {np.nonzero(wrong_in)[0][:5]}"
    assert not wrong_out.any(), f"filter certified OUTSIDE wrongly on rows {np.nonzero(wrong_out)[0][:5]}"
    assert np.all(blocks[:, 3] >= 0) and np.all(blocks[:, 4] > 0)
    decided = (verdict != 0).mean()
    return decided, verdict, want

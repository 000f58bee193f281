"""Shared predicate test material: exact signs with fractions.Fraction and adversarial (near-degenerate) inputs."""
from fractions import Fraction as F

import numpy as np


def _det(m):
    n = len(m)
    if n == 1:
        return m[0][0]
    if n == 2:
        return m[0][0] * m[1][1] - m[0][1] * m[1][0]
    s = F(0)
    for j in range(n):
        if m[0][j] == 0:
            continue
        minor = [row[:j] + row[j + 1:] for row in m[1:]]
        s += (-1) ** j * m[0][j] * _det(minor)
    return s


def _sgn(x):
    return (x > 0) - (x < 0)


def orient2d_exact(r):
    a, b, c = [[F(float(x)) for x in r[i:i + 2]] for i in (0, 2, 4)]
    return _sgn(_det([[a[0] - c[0], a[1] - c[1]], [b[0] - c[0], b[1] - c[1]]]))


def orient3d_exact(r):
    p = [[F(float(x)) for x in r[i:i + 3]] for i in (0, 3, 6, 9)]
    d = p[3]
    return _sgn(_det([[q[k] - d[k] for k in range(3)] for q in p[:3]]))


def incircle_exact(r):
    p = [[F(float(x)) for x in r[i:i + 2]] for i in (0, 2, 4, 6)]
    d = p[3]
    rows = []
    for q in p[:3]:
        x, y = q[0] - d[0], q[1] - d[1]
        rows.append([x, y, x * x + y * y])
    return _sgn(_det(rows))


def insphere_exact(r):
    p = [[F(float(x)) for x in r[i:i + 3]] for i in (0, 3, 6, 9, 12)]
    e = p[4]
    rows = []
    for q in p[:4]:
        x, y, z = q[0] - e[0], q[1] - e[1], q[2] - e[2]
        rows.append([x, y, z, x * x + y * y + z * z])
    # sign convention of Shewchuk's insphere: positive inside when orient3d(a,b,c,d) > 0
    return _sgn(_det(rows))


EXACT = {"orient2d": orient2d_exact, "orient3d": orient3d_exact, "incircle": incircle_exact, "insphere": insphere_exact}
WIDTH = {"orient2d": (3, 2), "orient3d": (4, 3), "incircle": (4, 2), "insphere": (5, 3)}


def _perturb(a, rng, max_ulps):
    k = rng.integers(-max_ulps, max_ulps + 1, size=a.shape)
    k = np.where(a == 0.0, 0, k)  # a perturbed zero is a denormal: 1000+ bits of dynamic range (ERR_RANGE by design)
    out = a.copy()
    for _ in range(max_ulps):
        up = k > 0
        dn = k < 0
        out = np.where(up, np.nextafter(out, np.inf), np.where(dn, np.nextafter(out, -np.inf), out))
        k = k - np.sign(k)
    return out


def adversarial(kind, n, seed=0):
    """n rows for predicate `kind`: a mix of random, exactly degenerate and few-ulp-perturbed degenerate inputs."""
    rng = np.random.default_rng(seed)
    npts, dim = WIDTH[kind]
    rows = []
    for i in range(n):
        mode = i % 4
        if mode == 0:  # generic
            pts = rng.random((npts, dim))
        else:
            # exactly degenerate configuration on a small integer lattice
            if kind in ("orient2d", "orient3d"):
                base = rng.integers(-8, 9, size=(npts - 1, dim)).astype(float)
                w = rng.integers(-3, 4, size=npts - 1).astype(float)
                w[-1] = 1 - w[:-1].sum()  # affine combination => collinear / coplanar
                last = (w[:, None] * base).sum(axis=0)
                pts = np.vstack([base, last])
            else:
                # points on a common circle / sphere: signed permutations of one integer vector
                v = np.array([1.0, 8.0, 4.0][:dim]) if dim == 3 else np.array([7.0, 4.0])
                pts = []
                while len(pts) < npts:
                    perm = rng.permutation(dim)
                    sg = rng.choice([-1.0, 1.0], size=dim)
                    cand = (v[perm] * sg).tolist()
                    if cand not in pts:
                        pts.append(cand)
                pts = np.array(pts)
            scale = 2.0 ** rng.integers(-20, 4)
            shift = rng.integers(-4, 5, size=dim).astype(float) * scale * 8
            pts = pts * scale + shift
            if mode == 2:
                pts = _perturb(pts, rng, 1)
            elif mode == 3:
                pts = _perturb(pts, rng, 3)
        rows.append(pts.reshape(-1))
    return np.array(rows)


def wide_range(kind, n, seed=0, span=120):
    """rows whose coordinates span ~2^span in magnitude (exercises the limb budget of the integer path)."""
    rng = np.random.default_rng(seed)
    npts, dim = WIDTH[kind]
    e = rng.integers(-span // 2, span // 2, size=(n, npts * dim))
    m = rng.integers(1, 2 ** 20, size=(n, npts * dim)).astype(float) * rng.choice([-1.0, 1.0], size=(n, npts * dim))
    return m * (2.0 ** e)

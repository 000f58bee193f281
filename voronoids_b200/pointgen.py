"""Synthetic point sets: counter-based splitmix64, identical in
numpy (here), C (oracle) and CUDA (csrc/pointgen.cu).

    x = splitmix64(seed * 0x9E3779B97F4A7C15 + counter)
    u = (x >> 11) * 2**-53                     in [0, 1)
coordinate j of point i uses counter i*N + j.

The reference's own tests draw from rand 0.8.5 StdRng (ChaCha12), which cannot
be reproduced without Rust; none of its assertions depend on the values
(/root/reference/tests/test_delaunay_tree.rs:9-10).
"""
import numpy as np

_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(v):
    """splitmix64 output function applied to the uint64 array `v` (state + gamma, then finalise)."""
    with np.errstate(over="ignore"):
        z = v.astype(np.uint64) + _GAMMA
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _u01(seed, counters):
    with np.errstate(over="ignore"):
        x = splitmix64(np.uint64(seed) * _GAMMA + counters.astype(np.uint64))
    return (x >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def uniform(n, dim=3, seed=0, first=0):
    """n points uniform in [0,1)^dim, float64 [n, dim]; `first` = index of the first point."""
    c = (np.arange(first * dim, (first + n) * dim, dtype=np.uint64)).reshape(n, dim)
    return _u01(seed, c)


def gaussian_mixture(n, dim=3, seed=1, centres=64, sigma=0.01):
    """Clustered set C4a: `centres` centres uniform in [0.1,0.9]^dim, sigma-Gaussian blobs (Box-Muller)."""
    cen = 0.1 + 0.8 * uniform(centres, dim, seed=seed + 7919)
    which = (_u01(seed + 1, np.arange(n, dtype=np.uint64)) * centres).astype(np.int64)
    cnt = np.arange(n * dim * 2, dtype=np.uint64).reshape(n, dim, 2)
    u1 = _u01(seed + 2, cnt[..., 0])
    u2 = _u01(seed + 2, cnt[..., 1])
    g = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
    return cen[which] + sigma * g


def jittered_lattice(n, dim=3, seed=2, jitter=1e-9):
    """Jittered lattice C4b: m^dim lattice (x fastest), spacing h=1/m, jitter uniform in +-jitter*h."""
    m = int(np.ceil(n ** (1.0 / dim) - 1e-9))
    while m ** dim < n:
        m += 1
    idx = np.arange(n, dtype=np.int64)
    coords = np.empty((n, dim), dtype=np.float64)
    h = 1.0 / m
    r = idx.copy()
    for k in range(dim):
        coords[:, k] = (r % m) * h
        r //= m
    u = uniform(n, dim, seed=seed)
    return coords + (2.0 * u - 1.0) * (jitter * h)


def make(kind, n, dim=3, seed=0):
    if kind == "uniform":
        return uniform(n, dim, seed)
    if kind == "clustered":
        return gaussian_mixture(n, dim, seed)
    if kind == "lattice":
        return jittered_lattice(n, dim, seed)
    raise ValueError(kind)


def uniform_sets_torch(n_sets, n, dim=3, first_seed=1000, device="cuda"):
    """`n_sets` uniform sets of `n` points generated ON THE DEVICE (BASELINE.json configs[4]: set s uses seed first_seed + s),
    bit-identical to uniform(n, dim, first_seed + s): the same splitmix64 in torch int64 arithmetic (two's complement
    products wrap like uint64; logical right shifts are emulated by masking).  float64 tensor [n_sets * n, dim]."""
    import torch

    def lsr(x, k):   # logical shift right of an int64 holding a uint64 bit pattern
        return (x >> k) & ((1 << (64 - k)) - 1)

    def i64(c):      # uint64 constant -> the int64 with the same bits
        return c - (1 << 64) if c >= (1 << 63) else c

    gamma = i64(0x9E3779B97F4A7C15)
    seeds = torch.arange(first_seed, first_seed + n_sets, dtype=torch.int64, device=device).view(n_sets, 1)
    ctr = torch.arange(n * dim, dtype=torch.int64, device=device).view(1, n * dim)
    z = seeds * gamma + ctr + gamma
    z = (z ^ lsr(z, 30)) * i64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * i64(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    return (lsr(z, 11).to(torch.float64) * (2.0 ** -53)).view(n_sets * n, dim)

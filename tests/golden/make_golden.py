"""Generates tests/golden/golden.json: canonical Delaunay edge lists (hashes) of the synthetic configurations,
computed in the build container by the exact oracle and cross-checked against scipy/Qhull on P u S where Qhull
finishes quickly.  The reference itself (Rust) cannot run here, see DESIGN.md "Oracle".

    python tests/golden/make_golden.py [--big] [--only name,name]   (--big adds the 10M / 5M-point cases, minutes / ~20 GB)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from voronoids_b200 import pointgen  # noqa: E402
from voronoids_b200._capi import edge_checksum_host  # noqa: E402

CASES = [
    # name, dim, kind, n, seed, qhull cross-check
    ("u3_10k", 3, "uniform", 10_000, 0, True),          # BASELINE.json configs[0]
    ("u3_100k", 3, "uniform", 100_000, 0, True),
    ("u3_1m", 3, "uniform", 1_000_000, 0, False),
    ("u2_10k", 2, "uniform", 10_000, 0, True),
    ("u2_1m", 2, "uniform", 1_000_000, 0, True),         # configs[1]
    ("c3_100k", 3, "clustered", 100_000, 1, True),       # configs[3] at n/50
    ("l3_100k", 3, "lattice", 100_000, 2, True),
    ("c3_500k", 3, "clustered", 500_000, 1, False),
    ("l3_500k", 3, "lattice", 500_000, 2, False),
    ("u3_set1000_100k", 3, "uniform", 100_000, 1000, False),  # configs[4]: set 0 of the batch
]
BIG = [("u3_10m", 3, "uniform", 10_000_000, 0, False),    # configs[2]
       ("c3_5m", 3, "clustered", 5_000_000, 1, False),    # configs[3] at full size
       ("l3_5m", 3, "lattice", 5_000_000, 2, False)]


def run(case):
    name, dim, kind, n, seed, qh = case
    p = pointgen.make(kind, n, dim, seed)
    t0 = time.time()
    ex = O.ExactDelaunay(p)
    e = ex.edges()
    rec = {"name": name, "dim": dim, "kind": kind, "n": n, "seed": seed, "n_edges": int(len(e)), "sha256": O.edge_sha256(e),
           "checksum64": edge_checksum_host(e), "live_simplices": ex.stats()["live"], "oracle_exact_calls": ex.pred["exact"],
           "oracle_exact_zero": ex.pred["zero"], "qhull_checked": False}
    if qh:
        from scipy.spatial import Delaunay
        sup = O.ref_super_simplex(p)[0]
        q = Delaunay(np.vstack([sup, p]))
        s1 = np.sort(q.simplices, axis=1)
        s1 = s1[np.lexsort(s1.T[::-1])]
        s2 = np.sort(ex.simplices(), axis=1)
        s2 = s2[np.lexsort(s2.T[::-1])]
        if kind == "uniform":
            assert np.array_equal(s1, s2), f"{name}: exact oracle and Qhull disagree"
            rec["qhull_checked"] = True
        else:
            # Qhull is a floating-point code: on clustered / near-degenerate input it may flip a few nearly
            # cospherical simplices.  The oracle is exact and self-validated (vo_bw_validate), so only record it.
            a, b = set(map(tuple, s1)), set(map(tuple, s2))
            rec["qhull_only_simplices"] = len(a - b)
            rec["oracle_only_simplices"] = len(b - a)
        assert ex.validate() == 0
    print(name, rec["n_edges"], rec["sha256"][:16], "%.1fs" % (time.time() - t0), flush=True)
    return rec


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    old = {}
    if os.path.exists(path):
        old = {r["name"]: r for r in json.load(open(path))["cases"]}
    cases = CASES + (BIG if "--big" in sys.argv else [])
    if "--only" in sys.argv:
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))
        cases = [c for c in CASES + BIG if c[0] in only]
    for c in cases:
        old[c[0]] = run(c)
    json.dump({"generator": "tests/golden/make_golden.py", "rng": "splitmix64 counter RNG, voronoids_b200/pointgen.py",
               "cases": list(old.values())}, open(path, "w"), indent=1)

"""Scratch (compute-sanitizer): MID twin on a jittered lattice, edges, validation, a batch tree, incremental inserts."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from voronoids_b200 import _capi, _lib, pointgen
lib = _lib.lib()
p = pointgen.make("lattice", 30000, 3, 2)
t = _capi.Tree(lib, p); e = t.edges(); ok = t.check_delaunay()[0]; st = t.stats(); t.close()
print("lattice", len(e), ok, st["flagged"])
sets = [pointgen.uniform(n, 3, 7 + n) for n in (3000, 40, 900)]
off = np.zeros(4, dtype=np.int64); off[1:] = np.cumsum([len(s) for s in sets])
t = _capi.Tree(lib, np.concatenate(sets), set_offsets=off); print("batch", len(t.edges()), t.check_delaunay()[0]); t.close()
q = pointgen.uniform(5000, 2, 3)
t = _capi.Tree(lib, q, insert=False)
for a, b in ((0, 1), (1, 300), (300, 5000)):
    t.insert(q[a:b], mode=1)
print("incremental", len(t.edges()), t.check_delaunay()[0]); t.close()

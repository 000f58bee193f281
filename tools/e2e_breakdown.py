"""Scratch: where does the end-to-end time go?  python tools/e2e_breakdown.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import voronoids_b200 as vb
from voronoids_b200 import pointgen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pts = pointgen.uniform(n, 3, 0)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tree = vb.delaunay(pts, device=0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    e = tree.edges()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    tree.close()
    t3 = time.perf_counter()
    print("iter %d: delaunay %.1f ms, edges %.1f ms (%d edges), close %.1f ms" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(e), (t3 - t2) * 1e3), flush=True)

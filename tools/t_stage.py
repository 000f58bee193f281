"""Scratch timing helper (not part of the product): python t_stage.py N DIM [stats] -- insert time per iteration, engine options from VOR_* env."""
import ctypes as C, numpy as np, time, torch, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from voronoids_b200 import _capi, _lib, pointgen
lib = _lib.lib()
n = int(sys.argv[1]); dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
stats = len(sys.argv) > 3
p = torch.from_numpy(pointgen.uniform(n, dim, 0)).cuda()
best = 1e9
for it in range(4):
    lib.vor_set_option(b"verbose", 1.0 if (it == 3 and os.environ.get("T_VERBOSE")) else 0.0)
    lib.vor_set_option(b"stats", 1.0 if (it == 3 and stats) else 0.0)
    lib.vor_set_option(b"profile", 1.0 if (it == 3 and os.environ.get("T_PROFILE")) else 0.0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h = _capi.tree_p()
    st = lib.vor_tree_create_device(dim, C.c_void_p(p.data_ptr()), n, 0, (C.c_void_p(torch.cuda.current_stream().cuda_stream) if os.environ.get('T_STREAM') else None), C.byref(h)); torch.cuda.synchronize(); t1 = time.perf_counter()
    st = lib.vor_tree_insert_device(h, C.c_void_p(p.data_ptr()), n, 1); torch.cuda.synchronize(); t2 = time.perf_counter()
    assert st == 0, st
    if it == 3 and stats:
        s = (C.c_uint64 * 32)()
        lib.vor_tree_stats(h, s)
        print("stats", list(s)[:20])
    if it == 3 and os.environ.get("T_PROFILE"):
        pr = (C.c_double * 8)()
        lib.vor_tree_profile(h, pr)
        print("PROFILE attempt_ms=%.2f commit_ms=%.2f setup_ms=%.2f total_ms=%.2f" % (pr[0], pr[2], pr[3], (t2 - t0) * 1e3))
    lib.vor_tree_destroy(h); torch.cuda.synchronize()
    if it and not (it == 3 and (stats or os.environ.get("T_VERBOSE") or os.environ.get("T_PROFILE"))): best = min(best, (t2 - t0) * 1e3)
print("RESULT n=%d dim=%d env=%s best_ms=%.2f Mpts/s=%.2f" % (n, dim, {k: v for k, v in os.environ.items() if k.startswith("VOR_")}, best, n / best / 1e3), flush=True)

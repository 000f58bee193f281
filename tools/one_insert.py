"""Scratch helper for ncu (not part of the product): ONE create + insert of N uniform points, engine options from VOR_* env."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from voronoids_b200 import _capi, _lib, pointgen
lib = _lib.lib()
n = int(sys.argv[1]); dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = torch.from_numpy(pointgen.uniform(n, dim, 0)).cuda()
h = _capi.tree_p()
assert lib.vor_tree_create_device(dim, C.c_void_p(p.data_ptr()), n, 0, None, C.byref(h)) == 0
assert lib.vor_tree_insert_device(h, C.c_void_p(p.data_ptr()), n, 1) == 0
torch.cuda.synchronize()
print("done", n, dim)

import ctypes as C, numpy as np, time, torch, sys
sys.path.insert(0,'/root/repo')
from voronoids_b200 import _capi, _lib, pointgen
lib = _lib.lib()
n = int(sys.argv[1]); dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = torch.from_numpy(pointgen.uniform(n,dim,0)).cuda()
for it in range(3):
    lib.vor_set_option(b"verbose", 1.0 if it==2 else 0.0)
    torch.cuda.synchronize(); t0=time.perf_counter()
    h = _capi.tree_p()
    st = lib.vor_tree_create_device(dim, C.c_void_p(p.data_ptr()), n, 0, None, C.byref(h)); torch.cuda.synchronize(); t1=time.perf_counter()
    st = lib.vor_tree_insert_device(h, C.c_void_p(p.data_ptr()), n, 1); torch.cuda.synchronize(); t2=time.perf_counter()
    lib.vor_tree_destroy(h); torch.cuda.synchronize(); t3=time.perf_counter()
    print("iter", it, "create %.1f ms insert %.1f ms destroy %.1f ms"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3), flush=True)

cat > /tmp/t.py <<'PY'
import ctypes as C, numpy as np, time, torch, sys
sys.path.insert(0,'/root/repo')
from voronoids_b200 import _capi, _lib, pointgen
lib = _lib.lib()
n = int(sys.argv[1])
p = torch.from_numpy(pointgen.uniform(n,3,0)).cuda()
for it in range(3):
    lib.vor_set_option(b"verbose", 1.0 if it==2 else 0.0)
    torch.cuda.synchronize(); t0=time.perf_counter()
    h = _capi.tree_p()
    st = lib.vor_tree_create_device(3, C.c_void_p(p.data_ptr()), n, 0, None, C.byref(h)); torch.cuda.synchronize(); t1=time.perf_counter()
    st = lib.vor_tree_insert_device(h, C.c_void_p(p.data_ptr()), n, 1); torch.cuda.synchronize(); t2=time.perf_counter()
    lib.vor_tree_destroy(h); torch.cuda.synchronize(); t3=time.perf_counter()
    print("iter", it, "create %.1f ms insert %.1f ms destroy %.1f ms"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3), flush=True)
PY
VOR_ATTEMPT_DIV=64 VOR_MIN_ATTEMPT=8192 python /tmp/t.py 10000000 2>&1 | tail -30
VOR_ATTEMPT_DIV=64 VOR_MIN_ATTEMPT=8192 python /tmp/t.py 1000000 2>&1 | tail -24
VOR_GROUP=8 VOR_ATTEMPT_DIV=64 VOR_MIN_ATTEMPT=8192 ncu --set full --clock-control none --import-source on -k regex:k_attempt_coop -s 800 -c 3 -o gpurun_out/prof_attempt_big_r1 python bench.py --workload u3_10m --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b4.log 2>&1; tail -2 gpurun_out/ncu_b4.log

"""Compiles tests/cpp/hpp_smoke.cpp against include/voronoids.hpp, links it to a library that exports the C ABI (the
kernel emulation on the CPU, the product library on the B200) and checks its output against the exact oracle."""
import os
import subprocess

import numpy as np

from voronoids_b200 import pointgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(tmp_path, oracle, so_path, n):
    exe = os.path.join(str(tmp_path), "hpp_smoke")
    libdir, libfile = os.path.split(os.path.abspath(so_path))
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "hpp_smoke.cpp"),
                           "-o", exe, "-L" + libdir, "-l:" + libfile, "-Wl,-rpath," + libdir])
    pts = pointgen.uniform(n, 3, 5)
    f = os.path.join(str(tmp_path), "pts.bin")
    pts.tofile(f)
    out = subprocess.run([exe, f, str(n)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    l1, l2 = out.stdout.strip().splitlines()[-2:]
    w = l1.split()
    e = oracle.ExactDelaunay(pts).edges().astype(np.uint64)
    assert w[0] == "max0" and int(w[1]) == 4
    assert int(w[3]) == 8 + n and int(w[5]) == 1
    assert int(w[7]) == len(e) and int(w[9]) == int(e[:, 0].sum()) and int(w[10]) == int(e[:, 1].sum())
    c = [float(x) for x in l2.split()[1:]]
    assert c == [0.5, 0.5, 0.5, 0.8660254037844386]

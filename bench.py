#!/usr/bin/env python
"""bench.py -- Delaunay points inserted per second (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload u3_10m|u3_1m|u2_1m|c3_5m|l3_5m|batch_100k]
    python bench.py --impl reference ...      # the reference's CPU algorithm (oracle/refcpu.cpp port) on the host cores

A step = one full pass of the hot path over one synthetic point set: DelaunayTree::new (bounding sphere + super
simplex) followed by the insertion of every point (BRIO ordering, then rounds of locate / conflict / reserve /
retriangulate) on the device.
  value   points/s with the points already resident in HBM when the timed region starts (CUDA events on the
          library's stream, max over ranks); N>1 shards independent sets, one per rank (weak scaling, no collective).
  e2e     the same through the public Python API with HOST buffers: voronoids_b200.delaunay(points) + tree.edges()
          (H2D of the points and D2H of the canonical edge list inside the timed region).
  roofline  the attempt kernel (locate + conflict + reservation): algorithmic bytes / its CUDA-event time.
  cpu_baseline  oracle/refcpu.cpp, the float restatement of kazewong/Voronoids, timed on this box's cores on a bounded
          sample of the same workload (rank 0, N=1 only).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dim, kind, n, seed, description)
    "u3_10m": (3, "uniform", 10_000_000, 0, "3D uniform random 10M points (BASELINE.json configs[2])"),
    "u3_1m": (3, "uniform", 1_000_000, 0, "3D uniform random 1M points"),
    "u3_100k": (3, "uniform", 100_000, 0, "3D uniform random 100k points"),
    "u3_10k": (3, "uniform", 10_000, 0, "3D uniform random 10,000 points (BASELINE.json configs[0])"),
    "u2_1m": (2, "uniform", 1_000_000, 0, "2D uniform random 1M points (configs[1])"),
    "c3_5m": (3, "clustered", 5_000_000, 1, "3D Gaussian mixture 5M points (configs[3])"),
    "l3_5m": (3, "lattice", 5_000_000, 2, "3D jittered lattice 5M points (configs[3])"),
    # batch of independent sets (configs[4]): 64 sets x 100k points per GPU in ONE device store; set s of rank r uses
    # seed 1000 + 64 r + s
    "b3_64x100k": (3, "batch", 6_400_000, 1000, "batch of 64 independent 3D uniform sets x 100k points per GPU (configs[4] shape)"),
    # BASELINE.json configs[4] at its stated size: 8,192 sets x 100k points, FIXED total split over the ranks in contiguous
    # blocks (strong scaling), every rank streaming its block through vor_delaunay_batch_stream in chunks of 128 sets;
    # set s uses seed 1000 + s whatever the rank count; STREAM_SETS overrides the set count (VOR_STREAM_SETS)
    "b3_8192x100k": (3, "stream", 819_200_000, 1000, "batch of 8,192 independent 3D uniform sets x 100k points (BASELINE.json configs[4]), sharded over the GPUs"),
}
STREAM_SETS = int(os.environ.get("VOR_STREAM_SETS", "8192"))
# ONE triangulation over the GPUs (voronoids_b200/slab.py): fixed 10M-point set, strong scaling
WORKLOADS["u3_10m_slab"] = (3, "slab", 10_000_000, 0, "3D uniform random 10M points, ONE triangulation in slabs over the GPUs (halo exchange, certified)")
WORKLOADS["u3_1m_slab"] = (3, "slab", 1_000_000, 0, "3D uniform random 1M points, ONE triangulation in slabs over the GPUs (halo exchange, certified)")
BATCH_SETS, BATCH_SIZE = 64, 100_000


def algorithmic_bytes(dim, K, Cn, W=1.0):
    """bytes per inserted point, SoA store, u32 ids; E = K + C distinct in-sphere tests."""
    N, M = dim, dim + 1
    E = K + Cn
    attempt = 8 * N + W * (8 * M + 8 * N * M) + E * (4 * M + 8 * N * M) + K * 4 * M
    retri = Cn * (8 * M) + Cn * 4 + K * 4
    return attempt, attempt + retri


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.first = 0

    def mark(self):
        """The timed region starts here.  The process is started BEFORE the warm-up steps: nvidia-smi takes a few hundred ms to
        come up (NVML initialisation holds driver locks), which used to fall into the timed region and doubled the time of
        workloads whose whole timed region is a few tens of ms."""
        self.first = len(self.rows)

    def start(self):
        if os.environ.get("VOR_NO_SAMPLER"):   # diagnostics only: a bench line without clocks is not a valid bench line
            return
        # NVML in this process (the library nvidia-smi itself reads these fields from): a poll is two or three cheap calls.  A
        # separate nvidia-smi process polling every 200 ms was measured to stretch the timed steps of the 10M-point run from
        # 107-111 ms to 110-126 ms (tools/r2_exp23.sh, profiles/r2_bench_u3_10m.json history); it stays as the fallback.
        if not os.environ.get("VOR_SAMPLER_SMI") and self._start_nvml():
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _start_nvml(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                u = str(torch.cuda.get_device_properties(self.index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + u) if not u.startswith("GPU-") else u)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)     # fail here, not in the thread
            self.nvml_stop = threading.Event()

            def loop():
                while not self.nvml_stop.is_set():
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
                        r = int(reasons_fn(h))
                        act = lambda bit: "Active" if (r & bit) else "Not Active"
                        # same column order as the nvidia-smi query: sm, max, power, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
                        self.rows.append([str(sm), str(mx), "0", act(0x8), act(0x40), act(0x20), act(0x4)])
                    except Exception:
                        pass
                    self.nvml_stop.wait(0.2)
            self.th = threading.Thread(target=loop, daemon=True)
            self.th.start()
            self.source = "nvml (in-process, every 200 ms)"
            return True
        except Exception:
            return False

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if getattr(self, "nvml_stop", None) is not None:
            self.nvml_stop.set()
            self.th.join(timeout=2)
        elif not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[self.first:] or self.rows[-1:]    # a timed region shorter than the 200 ms sampling period: the last sample before it
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": getattr(self, "source", "nvidia-smi -lms 200")}


def bench_config(desc, n, dim, world):
    """Same dict for both arms (the driver compares them)."""
    return {"workload": desc, "points_per_gpu": n, "dim": dim, "gpus": world,
            "parallelism": f"{world} independent point set(s), one per GPU, no collective",
            "timed_region": "DelaunayTree::new + insertion of every point; value: input resident in HBM, e2e: host buffers in, edge list out",
            "l2": "inputs larger than L2 (%.0f MB of coordinates, multi-GB simplex store)" % (n * dim * 8 / 1e6)}


def make_points(name, rank):
    from voronoids_b200 import pointgen
    dim, kind, n, seed, desc = WORKLOADS[name]
    if kind == "batch":
        sets = [pointgen.uniform(BATCH_SIZE, dim, seed + BATCH_SETS * rank + s) for s in range(BATCH_SETS)]
        return np.concatenate(sets, axis=0), dim, n, desc
    return pointgen.make(kind, n, dim, seed + 7919 * rank), dim, n, desc


def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed on.  Passed to the restatement EXPLICITLY, because
    torch.distributed.run exports OMP_NUM_THREADS=1 to its workers."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def reference_points(name, ns):
    from voronoids_b200 import pointgen
    dim, kind, n, seed, desc = WORKLOADS[name]
    if kind in ("batch", "stream"):
        return pointgen.uniform(min(ns, BATCH_SIZE), dim, seed)   # one set of the batch (sets are independent units)
    if kind == "uniform":
        return pointgen.uniform(ns, dim, seed)                    # counter-based generator: a prefix of the workload
    return pointgen.make(kind, n, dim, seed)[:ns]


def reference_once(pts, nthreads):
    """The reference's own CPU path on `pts` (oracle/refcpu.cpp, restatement of kazewong/Voronoids lib.rs:104-125):
    DelaunayTree::new, the first 100,000 points one by one (TreeUpdate::new + insert_point), the rest in ONE
    add_points_to_tree call (make_queue -> find_placement -> rounds of insert_points_parallel).  Phases timed apart."""
    from oracle import oracle as O
    L = O.lib()
    p, pp = O._d(pts)
    n, dim = p.shape
    n_seq = min(n, 100_000)
    t0 = time.perf_counter()
    h = L.vo_ref_create(dim, pp, n)
    e = L.vo_ref_insert_sequential(h, pp, n_seq, 0)
    t1 = time.perf_counter()
    if not e and n > n_seq:
        rest = np.ascontiguousarray(p[n_seq:])
        e = L.vo_ref_add_points_to_tree(h, rest.ctypes.data_as(C.POINTER(C.c_double)), n - n_seq, n_seq, nthreads)
    t2 = time.perf_counter()
    L.vo_ref_destroy(h)
    if e:
        raise RuntimeError(f"reference restatement failed with code {e} (the Rust original would panic here)")
    return {"n": n, "n_sequential": n_seq, "n_parallel": n - n_seq, "t_total": t2 - t0, "t_sequential": t1 - t0, "t_parallel": t2 - t1}


CPU_FULL_SAMPLE = {3: 1_000_000, 2: 1_000_000}   # examples/parallel_insert.rs:7-9 shape: 100k sequential + 900k parallel


def cpu_sample_record(name, nthreads, ns, with_one_core):
    """One run of the reference path on the first `ns` points of the workload -> the cpu_baseline object."""
    dim = WORKLOADS[name][0]
    r = reference_once(reference_points(name, ns), nthreads)
    rec = {"value": r["n"] / r["t_total"], "unit": "points/s", "cores": nthreads, "kind": "port",
           "sample": (f"first {r['n']} points of {name} through the restated lib.rs:104-125 path: {r['n_sequential']} sequential inserts "
                      f"({r['t_sequential']:.1f} s) + {r['n_parallel']} in one add_points_to_tree call ({r['t_parallel']:.1f} s), "
                      f"{nthreads} OpenMP threads passed explicitly"),
           "sequential_phase_pts_per_s": r["n_sequential"] / r["t_sequential"] if r["t_sequential"] > 0 else None,
           "parallel_phase_pts_per_s": (r["n_parallel"] / r["t_parallel"]) if r["n_parallel"] else None}
    if with_one_core:
        # the same path on ONE thread, on a smaller sample (100k sequential + 100k parallel)
        r1 = reference_once(reference_points(name, min(ns, 200_000)), 1)
        rec["one_core"] = {"value": r1["n"] / r1["t_total"], "unit": "points/s", "sample": f"first {r1['n']} points, 1 thread",
                           "parallel_phase_pts_per_s": (r1["n_parallel"] / r1["t_parallel"]) if r1["n_parallel"] else None}
    return rec


def reference_arm(name, steps, warmup):
    """bench.py --impl reference: K timed steps of the reference's CPU path, each on a bounded sample of the workload
    sized so that the whole run ends within a few minutes; W warm-up steps on a small sample (a CPU has nothing to warm
    but its thread pool and page cache)."""
    dim = WORKLOADS[name][0]
    nthreads = host_threads()
    budget_s = float(os.environ.get("VOR_REF_BUDGET_S", "420"))
    # calibration + warm-up: 150k points (100k sequential + 50k parallel)
    cal = reference_once(reference_points(name, 150_000), nthreads)
    for _ in range(max(0, warmup - 1)):
        reference_once(reference_points(name, 150_000), nthreads)
    r_seq = cal["n_sequential"] / max(cal["t_sequential"], 1e-9)
    r_par = (cal["n_parallel"] / max(cal["t_parallel"], 1e-9)) if cal["n_parallel"] else r_seq
    per_step = budget_s / max(steps, 1)
    ns = int(min(CPU_FULL_SAMPLE[dim], max(300_000, 100_000 + (per_step - 100_000 / r_seq) * r_par)))
    ns = min(ns, WORKLOADS[name][2] if WORKLOADS[name][1] not in ("batch", "stream") else BATCH_SIZE)
    pts = reference_points(name, ns)
    runs = [reference_once(pts, nthreads) for _ in range(steps)]
    t = sum(r["t_total"] for r in runs) / len(runs)
    tp = sum(r["t_parallel"] for r in runs) / len(runs)
    ts = sum(r["t_sequential"] for r in runs) / len(runs)
    n = runs[0]["n"]
    cb = {"value": n / t, "unit": "points/s", "cores": nthreads, "kind": "port",
          "sample": (f"each step: first {n} points of {name} through the restated lib.rs:104-125 path ({runs[0]['n_sequential']} sequential "
                     f"inserts, {ts:.1f} s + {runs[0]['n_parallel']} in one add_points_to_tree call, {tp:.1f} s), {nthreads} OpenMP threads "
                     f"passed explicitly; sample sized for {steps} steps in ~{budget_s:.0f} s; warm-up steps on 150k points"),
          "sequential_phase_pts_per_s": runs[0]["n_sequential"] / ts if ts > 0 else None,
          "parallel_phase_pts_per_s": (runs[0]["n_parallel"] / tp) if runs[0]["n_parallel"] and tp > 0 else None}
    return cb, t, n


def run_stream(args, name, rank, world, local_rank):
    """configs[4]: a fixed batch of STREAM_SETS sets x 100k points, contiguous blocks of sets per rank, no data-path collective."""
    import hashlib
    import torch
    import torch.distributed as dist
    import voronoids_b200 as vb
    from voronoids_b200 import _capi, _lib, pointgen, sharding
    lib = _lib.lib()
    dim, kind, _, seed, desc = WORKLOADS[name]
    n_total_sets = STREAM_SETS
    mine = sharding.shard_range(n_total_sets, world, rank)
    ns = len(mine)
    n_pts = ns * BATCH_SIZE
    # generated on the device, 512 sets at a time (same generator as pointgen.uniform, seed 1000 + s)
    pts_dev = torch.empty((max(n_pts, 1), dim), dtype=torch.float64, device="cuda")
    for lo in range(0, ns, 512):
        k = min(512, ns - lo)
        pts_dev[lo * BATCH_SIZE:(lo + k) * BATCH_SIZE] = pointgen.uniform_sets_torch(k, BATCH_SIZE, dim, seed + mine.start + lo)
    off = np.arange(ns + 1, dtype=np.int64) * BATCH_SIZE
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        ne, ck = _capi.delaunay_batch_stream(lib, int(pts_dev.data_ptr()), off, device=local_rank, dim=dim) if ns else (np.zeros(0, np.uint64),) * 2
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), ne, ck

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    launches0 = lib.vor_kernel_launches()
    barrier()
    sampler.mark()
    t_ms = []
    for _ in range(args.steps):
        ms, ne, ck = step_device()
        t_ms.append(ms)
    barrier()
    if os.environ.get("VOR_BENCH_VERBOSE"):
        print("step ms:", ["%.2f" % x for x in t_ms], file=sys.stderr)
    clocks = sampler.stop()
    launches = (lib.vor_kernel_launches() - launches0) // max(args.steps, 1)
    tot = torch.tensor([sum(t_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot.item()) / args.steps
    n_all = n_total_sets * BATCH_SIZE
    value = n_all / (ms_per_step * 1e-3)
    per_set = sharding.gather_per_set({mine.start + i: (int(ne[i]), int(ck[i])) for i in range(ns)}, n_total_sets)
    digest = hashlib.sha256(np.array(per_set, dtype=np.uint64).tobytes()).hexdigest()

    # end to end: host buffers in (pageable, staged through pinned chunks, next chunk's copy under the current chunk's rounds),
    # per-set results out
    e2e = None
    if not args.no_e2e:
        pts_host = pts_dev.cpu().numpy()
        def step_e2e():
            t0 = time.perf_counter()
            a, b = vb.delaunay_batch_stream(pts_host, off, device=local_rank) if ns else (None, None)
            torch.cuda.synchronize()
            return time.perf_counter() - t0, a, b
        step_e2e()
        barrier()
        dts = []
        for _ in range(max(1, min(args.steps, 2))):
            dt, a, b = step_e2e()
            dts.append(dt)
        barrier()
        assert ns == 0 or (np.array_equal(a, ne) and np.array_equal(b, ck))
        te = torch.tensor([sum(dts) / len(dts)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": n_all / float(te.item()), "unit": "points/s", "h2d_bytes_per_step": int(n_pts * dim * 8), "d2h_bytes_per_step": int(ns * 16),
               "api": "voronoids_b200.delaunay_batch_stream(points, set_offsets) -> per-set (n_edges, checksum64)", "edges": int(sum(x[0] for x in per_set))}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_sample_record(name, host_threads(), BATCH_SIZE, with_one_core=False)
    if rank == 0:
        cfg = bench_config(desc, n_all // world, dim, world)
        cfg.update({"sets_total": n_total_sets, "points_per_set": BATCH_SIZE, "sets_per_gpu": -(-n_total_sets // world), "chunk_sets": 128,
                    "parallelism": f"{n_total_sets} independent sets in contiguous blocks over {world} GPU(s); no data-path collective, per-set results gathered on the host",
                    "per_set_results_sha256": digest})
        line = {"metric": "delaunay_points_inserted_per_sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": None, "cpu_baseline": cpu_baseline, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_slab(args, name, rank, world, local_rank):
    """One triangulation in slabs over the ranks (strong scaling); the timed region is the whole slab pipeline on device-resident
    points: bounds, coarse sample, halo rounds, certification, this rank's part of the canonical edge list."""
    import hashlib
    import torch
    import torch.distributed as dist
    from voronoids_b200 import _lib, pointgen, slab
    lib = _lib.lib()
    dim, kind, n, seed, desc = WORKLOADS[name]
    allp = torch.from_numpy(pointgen.uniform(n, dim, seed)).cuda()
    mine, gidx = slab.partition_by_axis(allp, world, rank, axis=0)
    del allp
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        res = slab.delaunay_slab(lib, mine, gidx, device=local_rank, axis=0)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), res

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    launches0 = lib.vor_kernel_launches()
    barrier()
    sampler.mark()
    t_ms = []
    for _ in range(args.steps):
        ms, res = step()
        t_ms.append(ms)
    barrier()
    clocks = sampler.stop()
    launches = (lib.vor_kernel_launches() - launches0) // max(args.steps, 1)
    tot = torch.tensor([sum(t_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot.item()) / args.steps
    infos = [res.info]
    if world > 1:
        infos = [None] * world
        dist.all_gather_object(infos, res.info)
    full = slab.gather_edges(res.edges)
    if rank == 0:
        cfg = bench_config(desc, n // world, dim, world)
        cfg.update({"parallelism": f"1 point set in {world} slab(s) along x: shared 1/16 coarse sample, halo + hull shell exchanged peer to peer "
                                   f"(NCCL send/recv), certification on the cached circumspheres",
                    "halo_rows_received_per_rank": [i.get("halo_rows_received", 0) for i in infos],
                    "halo_bytes_received_per_rank": [i.get("halo_rows_received", 0) * (dim + 1) * 8 for i in infos],
                    "coarse_points": infos[0].get("coarse_points", 0), "tree_points_per_rank": [i.get("tree_points", 0) for i in infos],
                    "certification_rounds": max(i.get("rounds", 0) for i in infos),
                    "simplices_certified_by_peers": [int(i.get("peer_certified", 0)) for i in infos],
                    "edges": int(full.shape[0]), "edges_sha256": hashlib.sha256(np.ascontiguousarray(full).tobytes()).hexdigest()})
        line = {"metric": "delaunay_points_inserted_per_sec", "value": n / (ms_per_step * 1e-3), "unit": "points/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": None, "cpu_baseline": None, "e2e": None,
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=os.environ.get("VOR_BENCH_WORKLOAD", "u3_10m"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload
    dim, kind, n, seed, desc = WORKLOADS[name]

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, t, ns = reference_arm(name, args.steps, args.warmup)
        cfg = bench_config(desc, n, dim, args.gpus)
        cfg["reference_sample"] = f"each CPU step runs the first {ns} points of this workload (bounded sample, see cpu_baseline.sample)"
        line = {"impl": "reference", "metric": "delaunay_points_inserted_per_sec", "value": cb["value"], "unit": "points/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: voronoids_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line (NCCL prints its version there otherwise)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if kind == "stream":
        return run_stream(args, name, rank, world, local_rank)
    if kind == "slab":
        return run_slab(args, name, rank, world, local_rank)
    import voronoids_b200 as vb
    from voronoids_b200 import _capi, _lib
    lib = _lib.lib()

    pts_host, dim, n, desc = make_points(name, rank)
    pts_dev = torch.from_numpy(pts_host).cuda()
    # the rounds are launched on a stream of their own, not on the legacy default stream: programmatic dependent launches do
    # not overlap there (measured: 1M points 37 ms on the default stream, 28 ms on any other)
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.synchronize()
    sptr = C.c_void_p(stream.cuda_stream)
    boff = np.arange(BATCH_SETS + 1, dtype=np.int64) * BATCH_SIZE

    def step_device(stats=False, profile=False):
        """create + insert with device-resident input on torch's current stream; returns (ms, stats dict)."""
        lib.vor_set_option(b"stats", 1.0 if stats else 0.0)
        lib.vor_set_option(b"profile", 1.0 if profile else 0.0)
        h = _capi.tree_p()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        if kind == "batch":
            st = lib.vor_tree_create_batch_device(dim, C.c_void_p(pts_dev.data_ptr()), boff.ctypes.data_as(_capi.i64p), BATCH_SETS, local_rank, sptr,
                                                  C.byref(h))
        else:
            st = lib.vor_tree_create_device(dim, C.c_void_p(pts_dev.data_ptr()), n, local_rank, sptr, C.byref(h))
        if st != 0:
            raise RuntimeError(lib.vor_last_error().decode())
        if kind == "batch":
            st = lib.vor_tree_insert_batch_device(h, C.c_void_p(pts_dev.data_ptr()), boff.ctypes.data_as(_capi.i64p))
        else:
            st = lib.vor_tree_insert_device(h, C.c_void_p(pts_dev.data_ptr()), n, 1)
        if st not in (0, 3):
            raise RuntimeError(lib.vor_last_error().decode())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        s = (C.c_uint64 * _capi.N_STATS)()
        lib.vor_tree_stats(h, s)
        sd = dict(zip(_capi.STAT_NAMES, [int(x) for x in s]))
        prof = (C.c_double * 8)()
        lib.vor_tree_profile(h, prof)
        sd["profile_ms"] = {"attempt": prof[0], "commit": prof[2], "spheres": prof[1], "setup": prof[3]}
        sd["profile_launches"] = {"attempt": prof[4], "commit": prof[6]}
        lib.vor_tree_destroy(h)
        return ms, sd

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps (barrier + synchronize on both sides; time = max over ranks)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    launches0 = lib.vor_kernel_launches()
    barrier()
    sampler.mark()
    t_ms = []
    for _ in range(args.steps):
        ms, _ = step_device()
        t_ms.append(ms)
    barrier()
    if os.environ.get("VOR_BENCH_VERBOSE"):
        print("step ms:", ["%.2f" % x for x in t_ms], file=sys.stderr)
    clocks = sampler.stop()
    launches = (lib.vor_kernel_launches() - launches0) // max(args.steps, 1)
    tot = torch.tensor([sum(t_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot.item()) / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- two instrumented steps (not timed): per-kernel CUDA-event times (counters off, like the timed steps), then
    # the W/E/K/C counters (their atomics perturb the kernels, so they get a step of their own)
    _, sp = step_device(profile=True)
    _, sd = step_device(stats=True)
    sd["profile_ms"] = sp["profile_ms"]
    sd["profile_launches"] = sp["profile_launches"]
    K = sd["killed"] / max(sd["winners"], 1)
    Cn = sd["created"] / max(sd["winners"], 1)
    b_attempt, b_total = algorithmic_bytes(dim, K, Cn)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    t_attempt = sd["profile_ms"]["attempt"] * 1e-3
    achieved = (n * b_attempt / t_attempt / 1e9) if t_attempt > 0 else None
    # DRAM traffic of the kernel from the committed ncu --set full capture of the largest round (every slot of that launch holds a
    # live attempt: bytes per ATTEMPT), scaled to the attempts of the average launch of this run
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "attempt_traffic.json")))
        if tr.get("dim") == dim:
            traffic = tr["dram_bytes_per_attempt"] * sd["attempts"] / max(sd["profile_launches"]["attempt"], 1)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_attempt_hot + its exact twin k_attempt_slow (locate + conflict + reservation)", "achieved": achieved,
                "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "achieved_note": "algorithmic bytes of all launches (SURVEY.md 8d formula, measured K and C) / summed CUDA-event time of all "
                                 "launches (launch sizes vary by 4 orders of magnitude)",
                "algorithmic_bytes_per_point_kernel": b_attempt, "algorithmic_bytes_per_point_path": b_total,
                "algorithmic_bytes_per_launch": n * b_attempt / max(sd["profile_launches"]["attempt"], 1),
                "traffic_note": "dram__bytes read + write per attempt of the committed ncu --set full capture of the largest round (profiles/attempt_traffic.json) "
                                "x attempts per average launch of this run; compare with algorithmic_bytes_per_launch",
                "kernel_launches": sd["profile_launches"]["attempt"], "kernel_ms_total": sd["profile_ms"]["attempt"],
                "path_frac": (value / world) * b_total / (peak * 1e9),
                "counters_per_point": {"W_walk_steps_all_attempts": sd["walk_steps"] / n, "E_tests_all_attempts": sd["tests"] / n,
                                       "K_killed": K, "C_created": Cn, "attempts_per_point": sd["attempts"] / n,
                                       "rounds": sd["rounds"], "exact_calls": sd["exact_calls"], "exact_zero": sd["exact_zero"],
                                       "aborted_attempts": sd["aborted"] / n, "E_tests_completed_attempts": sd["tests_completed"] / n,
                                       "sphere_filter_undecided_tests": sd["sphere_undecided"], "points_via_exact_twin": sd["flagged"],
                                       "attempt_slots_launched": sd["slots"] / n},
                "step_ms_by_kernel": sd["profile_ms"]}

    # ---- end to end through the public API with host buffers (H2D + D2H inside the timed region)
    e2e = None
    lib.vor_set_option(b"stats", 0.0)      # the instrumented steps above must not leak their counters (atomics) into the end-to-end steps
    lib.vor_set_option(b"profile", 0.0)
    if not args.no_e2e:
        def step_e2e():
            t0 = time.perf_counter()
            if kind == "batch":
                tree = vb.delaunay_batch([pts_host[s * BATCH_SIZE:(s + 1) * BATCH_SIZE] for s in range(BATCH_SETS)], device=local_rank)
            else:
                tree = vb.delaunay(pts_host, device=local_rank)
            e = tree.edges()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            ne = len(e)
            tree.close()
            return dt, ne
        step_e2e()
        barrier()
        dts = []
        for _ in range(max(1, min(args.steps, 3))):
            dt, ne = step_e2e()
            dts.append(dt)
        barrier()
        te = torch.tensor([sum(dts) / len(dts)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n / float(te.item()), "unit": "points/s", "h2d_bytes_per_step": int(n * dim * 8), "d2h_bytes_per_step": int(ne * 8),
               "api": ("voronoids_b200.delaunay_batch(sets) + result.edges()" if kind == "batch" else "voronoids_b200.delaunay(points) + tree.edges()"),
               "edges": ne}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_sample_record(name, host_threads(), CPU_FULL_SAMPLE[dim], with_one_core=True)

    if rank == 0:
        line = {"metric": "delaunay_points_inserted_per_sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": bench_config(desc, n, dim, world),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

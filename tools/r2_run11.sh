#!/bin/bash
cd "$(dirname "$0")/.."
V=r2g
python bench.py > gpurun_out/${V}_bench_main.json 2> gpurun_out/${V}_bench_main.err
VOR_PINNED_RESULTS=0 python bench.py --no-cpu-baseline > gpurun_out/${V}_bench_main_pageable.json 2>&1
for f in gpurun_out/${V}_bench_main.json gpurun_out/${V}_bench_main_pageable.json; do grep '^{' $f | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]/1e6,2),"Mpts/s", round(d["ms_per_step"],2),"ms", "frac",round(r["frac"],4), r["step_ms_by_kernel"], "e2e", d["e2e"] and round(d["e2e"]["value"]/1e6,2), d["cpu_baseline"])'; done
tail -3 gpurun_out/${V}_bench_main.err
VOR_VERBOSE=1 python tools/e2e_breakdown.py 2>&1 | grep -E "edges:|iter" | tail -12
nproc; free -g | head -2

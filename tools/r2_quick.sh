#!/bin/bash
# quick GPU check of a build (run under gpurun): parity suite, then the headline bench lines
cd "$(dirname "$0")/.."
V=${1:-r2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${V}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${V}_gpu_tests.log
tail -3 gpurun_out/${V}_gpu_tests.log
python bench.py --no-cpu-baseline > gpurun_out/${V}_bench_main.json 2> gpurun_out/${V}_bench_main.err
python bench.py --workload u3_1m --no-cpu-baseline > gpurun_out/${V}_bench_u3_1m.json 2>&1
python bench.py --workload u2_1m --no-cpu-baseline > gpurun_out/${V}_bench_u2_1m.json 2>&1
for f in gpurun_out/${V}_bench_*.json; do echo "$f: $(grep '^{' $f | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]/1e6,2),"Mpts/s", round(d["ms_per_step"],2),"ms", "frac",round(r["frac"],4), r["step_ms_by_kernel"], {k:round(v,3) if isinstance(v,float) else v for k,v in r["counters_per_point"].items()}, "e2e", d["e2e"] and round(d["e2e"]["value"]/1e6,2))')"; done
tail -5 gpurun_out/${V}_bench_main.err

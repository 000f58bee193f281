#!/bin/bash
# round-2 evidence set at HEAD (run under gpurun, 1 GPU): bench lines of every workload, ncu launch list, ncu --set full of one
# full-size round (3D and 2D), e2e breakdown
cd "$(dirname "$0")/.."
V=${1:-r2}
O=gpurun_out
python bench.py > $O/${V}_bench_u3_10m.json 2> $O/${V}_bench_u3_10m.err
python bench.py --workload u3_1m --no-cpu-baseline > $O/${V}_bench_u3_1m.json 2>&1
python bench.py --workload u2_1m --no-cpu-baseline > $O/${V}_bench_u2_1m.json 2>&1
python bench.py --workload u3_100k --no-cpu-baseline > $O/${V}_bench_u3_100k.json 2>&1
python bench.py --workload c3_5m --no-cpu-baseline > $O/${V}_bench_c3_5m.json 2>&1
python bench.py --workload l3_5m --no-cpu-baseline > $O/${V}_bench_l3_5m.json 2>&1
python bench.py --workload b3_64x100k --no-cpu-baseline > $O/${V}_bench_b3_64x100k.json 2>&1
python bench.py --workload b3_8192x100k --steps 1 --warmup 1 --no-cpu-baseline > $O/${V}_bench_b3_8192x100k.json 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/${V}_bench_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${V}_launches_u3_10m.csv python tools/one_insert.py 10000000 3 > $O/${V}_ncu_l.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:k_attempt_hot|k_commit_coop|k_spheres|k_attempt_slow" -s 2400 -c 4 -o $O/${V}_round_u3_10m -f python tools/one_insert.py 10000000 3 > $O/${V}_ncu3.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:k_attempt_hot|k_commit_coop|k_spheres" -s 1500 -c 3 -o $O/${V}_round_u2_4m -f python tools/one_insert.py 4000000 2 > $O/${V}_ncu2.log 2>&1
VOR_VERBOSE=1 python tools/e2e_breakdown.py 2>&1 | grep -E "edges:|iter" | tail -6 > $O/${V}_e2e_breakdown.log
for f in $O/${V}_bench_*.json; do echo "$f: $(grep '^{' $f | tail -1 | cut -c1-160)"; done
tail -2 $O/${V}_ncu3.log $O/${V}_ncu2.log

#!/bin/bash
# round-end measurement set (run under gpurun): every workload's bench line + ncu launch list of the headline run
cd "$(dirname "$0")/.."
V=${1:-v6}
python bench.py > gpurun_out/bench_${V}_main.json 2> gpurun_out/bench_${V}_main.err
python bench.py --workload u2_1m --no-cpu-baseline > gpurun_out/bench_${V}_u2_1m.json 2>&1
python bench.py --workload u3_1m --no-cpu-baseline > gpurun_out/bench_${V}_u3_1m.json 2>&1
python bench.py --workload c3_5m --no-cpu-baseline --no-e2e > gpurun_out/bench_${V}_c3_5m.json 2>&1
python bench.py --workload l3_5m --no-cpu-baseline --no-e2e > gpurun_out/bench_${V}_l3_5m.json 2>&1
python bench.py --workload b3_64x100k --no-cpu-baseline --no-e2e > gpurun_out/bench_${V}_b3_64x100k.json 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${V}_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_${V}.csv python tools/one_insert.py 10000000 3 > gpurun_out/ncu_l.log 2>&1
VOR_VERBOSE=1 python tools/e2e_breakdown.py 2>&1 | grep -E "edges:|iter" | tail -12 > gpurun_out/e2e_breakdown_${V}.log
for f in gpurun_out/bench_${V}_*.json; do echo "$f: $(grep '^{' $f | tail -1 | cut -c1-220)"; done
cat gpurun_out/e2e_breakdown_${V}.log

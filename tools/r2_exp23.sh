#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp23.log; : > $O
for w in u3_1m u2_1m; do
for e in A=1 VOR_NO_SAMPLER=1; do
echo "## $w $e" >> $O
env $e VOR_BENCH_VERBOSE=1 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "step ms" >> $O
done; done
echo "## u3_10m" >> $O
VOR_BENCH_VERBOSE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "step ms" >> $O
VOR_NO_SAMPLER=1 VOR_BENCH_VERBOSE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "step ms" >> $O
cat $O

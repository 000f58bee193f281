#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp30.log; : > $O
for e in A=1 VOR_NO_SAMPLER=1 A=2; do
echo "## $e" >> $O
env $e VOR_STREAM_SETS=1024 VOR_BENCH_VERBOSE=1 python bench.py --workload b3_8192x100k --steps 3 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep -E "step ms|^\{" | cut -c1-230 >> $O
done
cat $O

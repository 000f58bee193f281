#!/bin/bash
# what the clock sampler costs the timed steps: none / NVML in-process / nvidia-smi process
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp38.log; : > $O
for e in VOR_NO_SAMPLER=1 A=1 VOR_SAMPLER_SMI=1 VOR_NO_SAMPLER=1 A=1 VOR_SAMPLER_SMI=1; do
echo "## $e" >> $O
env $e VOR_BENCH_VERBOSE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "step ms" >> $O
done
cat $O

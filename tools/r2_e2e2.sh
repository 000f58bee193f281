#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_e2e2.log; : > $O
VOR_VERBOSE=1 python tools/e2e_breakdown.py 10000000 >> $O 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline >> $O 2>&1
grep -v "stage \[" $O | tail -120

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_stage_times.log; : > $O
for a in "10000000 3" "1000000 3" "1000000 2"; do
echo "## t_stage.py $a" >> $O
T_VERBOSE=1 python tools/t_stage.py $a 2>&1 | grep -E "stage [0-9]+:|insert:|RESULT" >> $O
done
cat $O

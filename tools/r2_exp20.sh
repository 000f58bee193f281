#!/bin/bash
# stragglers handed to the next stage (carry_frac) + 2D lane groups of 16
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp20.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
for f in 0 0.03125 0.0625 0.125 0.25; do
run 100000 3 VOR_CARRY_FRAC=$f
run 1000000 3 VOR_CARRY_FRAC=$f
run 1000000 2 VOR_CARRY_FRAC=$f
run 10000000 3 VOR_CARRY_FRAC=$f
done
echo "## parity" >> $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
VOR_CARRY_FRAC=0.25 timeout 600 python -m pytest tests -m gpu -x -q -k "matches_oracle or incremental or golden or overflow or tiny or batch" 2>&1 | tail -3 >> $O
T_VERBOSE=1 python tools/t_stage.py 1000000 3 2>&1 | grep "stage [0-9]*:" >> $O
cat $O

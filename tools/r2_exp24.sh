#!/bin/bash
# per-entry flood-start hints (attempt prologue one round trip shorter)
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp24.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run 10000000 3 T_PROFILE=1
run 1000000 3 T_PROFILE=1
run 1000000 2 T_PROFILE=1
run 100000 3 A=1
run 8000000 2 A=1
echo "## parity" >> $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
cat $O

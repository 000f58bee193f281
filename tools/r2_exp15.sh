#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp15.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run 10000000 3 X=ridgehash
run 10000000 3 VOR_SO=variants/rh0.so
run 1000000 2 X=ridgehash
run 1000000 2 VOR_SO=variants/rh0.so
run 1000000 3 X=ridgehash
run 1000000 3 VOR_SO=variants/rh0.so
echo "## parity" >> $O
timeout 900 python -m pytest tests -m gpu -x -q -k "matches_oracle or options or batch_of or incremental or golden or overflow" 2>&1 | tail -3 >> $O
cat $O

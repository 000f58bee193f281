#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp7.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
for ma in 16384 32768 65536 131072; do
run 1000000 3 VOR_MIN_ATTEMPT=$ma
run 10000000 3 VOR_MIN_ATTEMPT=$ma
run 1000000 2 VOR_MIN_ATTEMPT=$ma
done
run 1000000 3 VOR_MIN_ATTEMPT=32768 VOR_ATTEMPT_DIV=32
run 10000000 3 VOR_MIN_ATTEMPT=32768 VOR_ATTEMPT_DIV=32
run 100000 3 X=base
run 100000 3 VOR_MIN_ATTEMPT=32768
cat $O

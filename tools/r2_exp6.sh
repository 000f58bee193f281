#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp6.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
for n in 1000000 10000000; do
run $n 3 X=base
run $n 3 VOR_STAGE_LOG=2
run $n 3 VOR_STAGE_LOG=3
run $n 3 VOR_ROUNDS_PER_SYNC=4
run $n 3 VOR_ROUNDS_PER_SYNC=16
run $n 3 VOR_STAGE0=1024
run $n 3 VOR_MIN_ATTEMPT=4096
run $n 3 VOR_ATTEMPT_DIV=48
run $n 3 VOR_ATTEMPT_DIV=96
done
run 1000000 2 X=base
run 1000000 2 VOR_STAGE_LOG=2
run 1000000 2 VOR_ATTEMPT_DIV=32
run 1000000 2 VOR_ATTEMPT_DIV=128
cat $O

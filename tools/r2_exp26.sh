#!/bin/bash
# sweep of the round-loop knobs under stage hand-over
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp26.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
for e in A=1 VOR_ROUNDS_PER_SYNC=4 VOR_ROUNDS_PER_SYNC=6 VOR_ROUNDS_PER_SYNC=12 VOR_MIN_ATTEMPT=4096 VOR_MIN_ATTEMPT=16384 VOR_ATTEMPT_DIV=48 VOR_ATTEMPT_DIV=96 VOR_STAGE0=64 VOR_STAGE0=1024 VOR_COMPACT_FRAC=0.92 VOR_COMPACT_FRAC=0.75; do
run 100000 3 $e
run 1000000 3 $e
run 10000000 3 $e
run 1000000 2 $e
done
cat $O

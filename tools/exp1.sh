#!/bin/bash
# scratch A/B timing on the GPU box (not part of the product)
cd "$(dirname "$0")/.."
O=gpurun_out/exp1.log; : > $O
run() { echo "## $*" >> $O; env "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|Error|error|assert" >> $O; }
run A=0
run VOR_PREWALK=4
run VOR_PREWALK=1
run VOR_RED=1
run VOR_PREWALK=4 VOR_RED=1
run VOR_ROUNDS_PER_SYNC=16
run VOR_PREWALK=4 VOR_RED=1 VOR_ATTEMPT_DIV=48
export N=1000000 DIM=2
run A=0
run VOR_ATTEMPT_DIV=32
run VOR_ATTEMPT_DIV=16
run VOR_ATTEMPT_DIV=8
run VOR_ATTEMPT_DIV=16 VOR_MIN_ATTEMPT=4096
run VOR_ATTEMPT_DIV=16 VOR_MIN_ATTEMPT=16384
run VOR_ATTEMPT_DIV=16 VOR_PREWALK=4 VOR_RED=1
run VOR_ATTEMPT_DIV=16 VOR_ROUNDS_PER_SYNC=16
unset N DIM
echo "## verbose baseline" >> $O
T_VERBOSE=1 python tools/t_stage.py 10000000 3 2>&1 | grep -E "stage [0-9]+:|setup" | tail -24 >> $O
echo "## tests with PREWALK=4 RED=1" >> $O
VOR_PREWALK=4 VOR_RED=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
cat $O

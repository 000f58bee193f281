#!/bin/bash
# scratch A/B timing on the GPU box (not part of the product)
cd "$(dirname "$0")/.."
O=gpurun_out/exp4.log; : > $O
run() { echo "## $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|Error|error|assert|rror" >> $O; }
echo "## tests RED=1 (recycle default on)" >> $O
VOR_RED=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O
run VOR_RED=1
run VOR_RED=1 VOR_RECYCLE=0
run VOR_RED=1 VOR_SO=$PWD/variants/r80.so
run VOR_RED=1 VOR_SO=$PWD/variants/r80.so VOR_RECYCLE=0
run VOR_RED=1 VOR_ATTEMPT_DIV=48
export N=1000000 DIM=2
run VOR_RED=1
run VOR_RED=1 VOR_RECYCLE=0
cat $O

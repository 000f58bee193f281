#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/exp5.log; : > $O
run() { echo "## $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|Error|error|assert|rror" >> $O; }
run VOR_RED=1
run VOR_RED=1 VOR_RECYCLE=1
run VOR_RED=1 VOR_COMMIT_SMEM=0
run VOR_RED=0
run VOR_RED=1 VOR_PERSIST=16384
run VOR_RED=1 VOR_PERSIST=16384 VOR_RECYCLE=1
echo "## tests RED=1 PERSIST=1024 RECYCLE=1" >> $O
VOR_RED=1 VOR_PERSIST=1024 VOR_RECYCLE=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
export N=1000000 DIM=2
run VOR_RED=1
unset N DIM
echo "## tests RED=1" >> $O
VOR_RED=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
cat $O

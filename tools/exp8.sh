#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/exp8.log; : > $O
run() { echo "## $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|Error|error|assert|rror" >> $O; }
run A=0
run VOR_SO=$PWD/variants/dedup0.so
run A=0
run VOR_SO=$PWD/variants/dedup0.so
export N=1000000 DIM=2
run A=0
run VOR_SO=$PWD/variants/dedup0.so
unset N DIM
echo "## tests" >> $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
VOR_RED=0 timeout 600 python -m pytest tests -m gpu -x -q -k "oracle or golden" 2>&1 | tail -3 >> $O
cat $O

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/exp7.log; : > $O
run() { echo "## $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|Error|error|assert|rror" >> $O; }
run A=0
run VOR_SO=$PWD/variants/ab64.so
run VOR_SO=$PWD/variants/ab32r80.so
run A=0
run VOR_SO=$PWD/variants/ab64.so
export N=1000000 DIM=2
run A=0
run VOR_SO=$PWD/variants/ab64.so
cat $O

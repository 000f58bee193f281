#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/exp6.log; : > $O
run() { echo "## $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|Error|error|assert|rror" >> $O; }
for v in s1f1 s0f1 s1f0 s0f0; do run VOR_RED=1 VOR_SO=$PWD/variants/$v.so; done
for v in s1f1 s0f0; do run VOR_RED=0 VOR_SO=$PWD/variants/$v.so; done
for v in s1f1 s0f1 s1f0 s0f0; do run VOR_RED=1 VOR_SO=$PWD/variants/$v.so; done
cat $O

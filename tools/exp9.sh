#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/exp9.log; : > $O
run() { echo "## $*" >> $O; env "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|rror" >> $O; }
for d in 40 48 56 64 72 80 96; do run VOR_ATTEMPT_DIV=$d; done
run VOR_MIN_ATTEMPT=4096
run VOR_MIN_ATTEMPT=16384
run VOR_STAGE0=1024
run VOR_STAGE0=64
export N=5000000
run A=0
run VOR_ATTEMPT_DIV=48
run VOR_ATTEMPT_DIV=80
cat $O

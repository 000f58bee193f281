#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp5.log; : > $O
run() { echo "## $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py 10000000 3 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run X=base
run VOR_SMEM_PAD=9000
run VOR_SMEM_PAD=20000
run VOR_SMEM_PAD=40000
run VOR_SO=variants/i1h48.so
cat $O

#!/bin/bash
# scratch A/B timing on the GPU box (not part of the product)
cd "$(dirname "$0")/.."
O=gpurun_out/exp2.log; : > $O
run() { echo "## $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|Error|error|assert" >> $O; }
run A=0
run VOR_RED=1
run VOR_COMMIT_SMEM=0
run VOR_COMMIT_SMEM=0 VOR_RED=1
export N=1000000 DIM=2
run A=0
run VOR_RED=1
run VOR_COMMIT_SMEM=0
unset N DIM
echo "## tests default" >> $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
echo "## tests RED=1" >> $O
VOR_RED=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
cat $O

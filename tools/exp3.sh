#!/bin/bash
# scratch A/B timing of library variants on the GPU box (not part of the product)
cd "$(dirname "$0")/.."
O=gpurun_out/exp3.log; : > $O
run() { echo "## $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|Error|error|assert" >> $O; }
for v in variants/*.so; do
  run VOR_SO=$PWD/$v VOR_RED=1
done
run VOR_SO=$PWD/variants/r64_b64.so VOR_RED=0
export N=1000000 DIM=2
for v in variants/r64_b64.so variants/r96_b64.so variants/r64_b32.so; do
  run VOR_SO=$PWD/$v VOR_RED=1
done
cat $O

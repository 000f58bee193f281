#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp2.log; : > $O
echo "## parity of base" >> $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O
run() { echo "## $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py 10000000 3 stats 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run X=base
run VOR_SO=variants/h48b64.so
run VOR_SO=variants/h56b64.so
run VOR_SO=variants/h64b32.so
run VOR_SO=variants/h56b64.so VOR_COMPACT_FRAC=0.85
cat $O

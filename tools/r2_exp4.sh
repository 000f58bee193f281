#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp4.log; : > $O
run() { echo "## $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py 10000000 3 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run X=persist1
run VOR_PERSIST_WAVES=100000
run VOR_PERSIST_WAVES=2
run VOR_PERSIST_WAVES=4
run X=persist1_again
echo "## parity" >> $O
timeout 900 python -m pytest tests -m gpu -x -q -k "matches_oracle or options or batch or incremental" 2>&1 | tail -3 >> $O
cat $O

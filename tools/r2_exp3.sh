#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp3.log; : > $O
run() { echo "## $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py 10000000 3 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run X=base_i1_h40
run VOR_SO=variants/i1h56.so
run VOR_SO=variants/i0h56.so
run VOR_SO=variants/i1h48.so
run VOR_SO=variants/i1h56s3.so
run VOR_SO=variants/i1h56spec0.so
run VOR_SO=variants/i1h56.so VOR_COMPACT_FRAC=0.85
run VOR_SO=variants/i1h56.so VOR_COMPACT_FRAC=0.85 VOR_ROUNDS_PER_SYNC=4
echo "## parity of i1h56" >> $O
VOR_SO=variants/i1h56.so timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
cat $O

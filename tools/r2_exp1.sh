#!/bin/bash
# experiment: attempt-kernel variants on the 10M-point run (per-kernel CUDA-event times)
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp1.log; : > $O
run() { echo "## $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py 10000000 3 2>&1 | grep -E "PROFILE|RESULT|rror" >> $O; }
run X=base
run VOR_COMPACT_FRAC=0.85
run VOR_SO=variants/spec0.so
run VOR_SO=variants/r64b32.so
run VOR_SO=variants/r72b32.so
run VOR_SO=variants/r48b64.so
run VOR_SO=variants/r40b64.so
run VOR_SO=variants/r40b64.so VOR_COMPACT_FRAC=0.85
echo "## parity of base" >> $O
timeout 600 python -m pytest tests -m gpu -x -q -k "matches_oracle or options or sphere" 2>&1 | tail -3 >> $O
cat $O

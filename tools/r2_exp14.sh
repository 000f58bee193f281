#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp14.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run 10000000 3 X=base
run 10000000 3 VOR_SUBROUND=16384
run 10000000 3 VOR_SUBROUND=32768
run 10000000 3 VOR_SUBROUND=65536
run 10000000 3 VOR_SUBROUND=100000
echo "## parity subround" >> $O
VOR_SUBROUND=16384 timeout 900 python -m pytest tests -m gpu -x -q -k "matches_oracle or golden" 2>&1 | tail -3 >> $O
cat $O

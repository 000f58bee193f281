#!/bin/bash
# lane groups of the hot / commit kernels: variants built by tools/build_variant.sh (base, vA, vB, vC)
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp19.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
for v in base vA vB vC; do
run 10000000 3 VOR_SO=$PWD/variants/$v.so
run 1000000 3 VOR_SO=$PWD/variants/$v.so
run 1000000 2 VOR_SO=$PWD/variants/$v.so
run 8000000 2 VOR_SO=$PWD/variants/$v.so
done
for v in vA vB vC; do
echo "## parity $v" >> $O
VOR_SO=$PWD/variants/$v.so timeout 600 python -m pytest tests -m gpu -x -q -k "matches_oracle or incremental or golden or overflow or tiny" 2>&1 | tail -3 >> $O
done
cat $O

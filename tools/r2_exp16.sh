#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp16.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
for n in 10000000 1000000 100000; do
run $n 3 VOR_PDL=1
run $n 3 VOR_PDL=0
done
run 1000000 2 VOR_PDL=1
run 1000000 2 VOR_PDL=0
echo "## parity" >> $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
cat $O

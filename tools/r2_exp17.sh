#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp17.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
for p in 0 8192 16384 32768 65536 131072; do
run 10000000 3 VOR_PDL=$p
done
for p in 8192 32768; do
run 1000000 3 VOR_PDL=$p
run 1000000 2 VOR_PDL=$p
done
cat $O

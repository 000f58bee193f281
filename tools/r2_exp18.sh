#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp18.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
for f in 0 2048 4096 8192 16384; do
run 100000 3 VOR_FUSE_SPHERES=$f
run 1000000 3 VOR_FUSE_SPHERES=$f
run 1000000 2 VOR_FUSE_SPHERES=$f
done
run 10000000 3 VOR_FUSE_SPHERES=0
run 10000000 3 VOR_FUSE_SPHERES=8192
echo "## parity" >> $O
VOR_FUSE_SPHERES=8192 timeout 900 python -m pytest tests -m gpu -x -q -k "matches_oracle or options or batch_of or incremental or golden or overflow or tiny" 2>&1 | tail -3 >> $O
cat $O

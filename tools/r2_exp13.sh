#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp13.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" T_PROFILE=1 python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run 10000000 3 X=tiled
run 10000000 3 VOR_TILED=0
run 1000000 3 X=tiled
run 1000000 3 VOR_TILED=0
run 1000000 2 X=tiled
run 1000000 2 VOR_TILED=0
run 100000 3 X=tiled
run 100000 3 VOR_TILED=0
echo "## parity" >> $O
timeout 900 python -m pytest tests -m gpu -x -q -k "matches_oracle or options or batch_of or incremental or golden" 2>&1 | tail -3 >> $O
for w in u3_100k u3_1m; do
for e in "X=1" "VOR_NO_SAMPLER=1"; do
env $e python bench.py --workload $w --no-cpu-baseline --no-e2e | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("'$w' '$e'", round(d["ms_per_step"],2),"ms", d["clocks"])' >> $O
done; done
cat $O

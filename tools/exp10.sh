#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/exp10.log; : > $O
run() { echo "## N=$N $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|rror" >> $O; }
for N in 100000 1000000; do export N
  for b in 0 4096 16384 65536 1000000000; do run VOR_STAGE_BELOW=$b; done
done
export N=10000000
for b in 0 8192 16384 32768; do run VOR_STAGE_BELOW=$b; done
cat $O

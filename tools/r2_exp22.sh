#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp22.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run 1000000 2 A=1
run 1000000 2 T_STREAM=1
run 1000000 2 A=1
run 1000000 3 T_STREAM=1
for w in u3_1m u2_1m u3_100k u3_10k; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])" >> $O; done
cat $O

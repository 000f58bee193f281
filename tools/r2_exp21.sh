#!/bin/bash
# slotInfo packing (commit starts from one load), H2D copy threads, full GPU suite, bench lines
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp21.log; : > $O
run() { n=$1; d=$2; shift 2; echo "## n=$n d=$d $*" >> $O; env "$@" python tools/t_stage.py $n $d 2>&1 | grep -E "PROFILE|RESULT|rror|stats" >> $O; }
run 10000000 3 T_PROFILE=1
run 1000000 3 T_PROFILE=1
run 1000000 2 T_PROFILE=1
run 100000 3 A=1
run 8000000 2 A=1
echo "## parity" >> $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
echo "## bench" >> $O
python bench.py --steps 5 --warmup 3 --no-cpu-baseline >> $O 2>&1
VOR_COPY_THREADS=4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('copy_threads=4 e2e', d['e2e']['value'])" >> $O
for w in u3_1m u2_1m u3_100k u3_10k; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$w', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])" >> $O; done
cat $O

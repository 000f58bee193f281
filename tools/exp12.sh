#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/exp12.log; : > $O
run() { echo "## N=$N $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|rror" >> $O; }
export N=10000000
run A=0
run A=0
unset N
for w in l3_5m c3_5m u2_1m; do echo "## bench $w" >> $O; python bench.py --workload $w --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d[\"roofline\"]; print(d[\"value\"]/1e6, d[\"ms_per_step\"], r[\"step_ms_by_kernel\"], r[\"counters_per_point\"][\"exact_calls\"])" >> $O; done
echo "## tests" >> $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
cat $O

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/exp11.log; : > $O
run() { echo "## N=$N $*" >> $O; env T_PROFILE=1 "$@" python tools/t_stage.py ${N:-10000000} ${DIM:-3} 2>&1 | grep -E "RESULT|PROFILE|rror" >> $O; }
export N=10000000
run A=0
run VOR_SO=$PWD/variants/dd0.so
run A=0
run VOR_SO=$PWD/variants/dd0.so
export N=1000000 DIM=2
run A=0
run VOR_SO=$PWD/variants/dd0.so
unset N DIM
for w in l3_5m; do for so in "" "$PWD/variants/dd0.so"; do echo "## bench $w so=$so" >> $O; VOR_SO=${so:-$PWD/voronoids_b200/libvoronoids_b200.so} python bench.py --workload $w --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d[\"roofline\"]; print(d[\"value\"]/1e6, d[\"ms_per_step\"], r[\"step_ms_by_kernel\"], r[\"counters_per_point\"][\"exact_calls\"])" >> $O; done; done
cat $O

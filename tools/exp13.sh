#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/exp13.log; : > $O
echo "## tests" >> $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
for i in 1 2; do T_PROFILE=1 python tools/t_stage.py 10000000 3 | grep -E "PROFILE|RESULT" >> $O; done
T_PROFILE=1 VOR_SPLIT_EXACT=0 python tools/t_stage.py 10000000 3 | grep -E "PROFILE|RESULT" >> $O
for w in l3_5m c3_5m u2_1m; do echo "## bench $w" >> $O; python bench.py --workload $w --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d[\"roofline\"]; print(d[\"value\"]/1e6, d[\"ms_per_step\"], r[\"step_ms_by_kernel\"], r[\"counters_per_point\"][\"exact_calls\"])" >> $O; done
cat $O

#!/bin/bash
# double-double stage (VOR_DD=1) now that only the exact twin and the MID twin have predicates in their call trees
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp31.log; : > $O
bw() { w=$1; python bench.py --workload $w --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; c=r['counters_per_point']
print('$w', round(d['ms_per_step'],1), 'ms', r['step_ms_by_kernel'], 'flagged', c['points_via_exact_twin'], 'exact', c['exact_calls'], 'rounds', c['rounds'])" >> $O; }
for so in voronoids_b200/libvoronoids_b200.so variants/dd1.so; do
export VOR_SO=$PWD/$so
echo "## $so" >> $O
bw l3_5m; bw c3_5m; bw u3_10m
done
export VOR_SO=$PWD/variants/dd1.so
echo "## parity dd1" >> $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
cat $O

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp29.log; : > $O
bw() { w=$1; python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; c=r['counters_per_point']
print('$w', round(d['ms_per_step'],1), 'ms', r['step_ms_by_kernel'], 'flagged', c['points_via_exact_twin'], 'exact', c['exact_calls'], 'undecided', c['sphere_filter_undecided_tests'], 'rounds', c['rounds'])" >> $O; }
for so in voronoids_b200/libvoronoids_b200.so variants/mid3.so; do
export VOR_SO=$PWD/$so
echo "## $so" >> $O
bw l3_5m; bw l3_5m
done
export VOR_SO=$PWD/variants/mid3.so
echo "## parity mid3" >> $O
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or lattice or l3 or adversarial or matches_oracle" 2>&1 | tail -3 >> $O
cat $O

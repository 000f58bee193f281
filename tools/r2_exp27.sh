#!/bin/bash
# FP64 determinant middle stage inside the hot kernel (variants/mid.so) against HEAD
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp27.log; : > $O
for so in voronoids_b200/libvoronoids_b200.so variants/mid.so; do
export VOR_SO=$PWD/$so
echo "## $so" >> $O
python tools/t_stage.py 10000000 3 2>&1 | grep RESULT >> $O
T_PROFILE=1 python tools/t_stage.py 10000000 3 2>&1 | grep PROFILE >> $O
for w in l3_5m c3_5m; do python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; c=r['counters_per_point']
print('$w', round(d['ms_per_step'],1), 'ms', r['step_ms_by_kernel'], 'flagged', c['points_via_exact_twin'], 'exact', c['exact_calls'], 'undecided', c['sphere_filter_undecided_tests'], 'rounds', c['rounds'])" >> $O; done
done
echo "## parity mid" >> $O
VOR_SO=$PWD/variants/mid.so timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
cat $O

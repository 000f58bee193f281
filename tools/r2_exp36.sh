#!/bin/bash
# 3D edge list: wedge test against pivots around every edge
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp36.log; : > $O
for w in 1 0; do
echo "## VOR_EDGE_WEDGE=$w" >> $O
VOR_EDGE_WEDGE=$w VOR_VERBOSE=1 python tools/e2e_breakdown.py 2>&1 | grep -E "edges:|iter" | tail -5 >> $O
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])" >> $O
cat $O

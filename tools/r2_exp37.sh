#!/bin/bash
# 3D edge list: wedge count pass as a kernel of its own against pivots around every edge
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp37.log; : > $O
for w in 1 0; do
echo "## VOR_EDGE_WEDGE=$w" >> $O
VOR_EDGE_WEDGE=$w VOR_VERBOSE=1 python tools/e2e_breakdown.py 2>&1 | grep -E "edges:|iter" | tail -5 >> $O
done
timeout 600 python -m pytest tests -m gpu -x -q -k "wedge or golden or matches_oracle or batch or slab" 2>&1 | tail -2 >> $O
cat $O

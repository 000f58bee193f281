#!/bin/bash
# last check at HEAD on one GPU: smoke, the whole GPU suite, the default bench line
cd "$(dirname "$0")/.."
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $O/r2b_bench_u3_10m.json 2> $O/r2b_bench_u3_10m.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2b_bench_reference.json 2>&1
grep '^{' $O/r2b_bench_u3_10m.json | tail -1 | cut -c1-400
tail -n 3 $O/r2b_bench_u3_10m.err

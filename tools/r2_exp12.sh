#!/bin/bash
cd "$(dirname "$0")/.."
for w in u3_100k u3_1m; do
for e in "X=1" "VOR_NO_SAMPLER=1"; do
env $e python bench.py --workload $w --no-cpu-baseline --no-e2e | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("'$w' '$e'", round(d["ms_per_step"],2),"ms", d["clocks"])'
done; done
T_STREAM=1 python tools/t_stage.py 100000 3 | tail -1
python tools/t_stage.py 100000 3 | tail -1
VOR_STREAM_SETS=1024 VOR_NO_SAMPLER=1 python bench.py --workload b3_8192x100k --steps 2 --warmup 1 --no-cpu-baseline --no-e2e | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("stream1024 nosampler", round(d["ms_per_step"],2),"ms")'
VOR_STREAM_SETS=2048 python bench.py --workload b3_8192x100k --steps 1 --warmup 1 --no-cpu-baseline --no-e2e | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("stream2048 sampler", round(d["ms_per_step"],2),"ms")'

#!/bin/bash
# final round-2 evidence set at HEAD (run under gpurun, 1 GPU): bench lines of every workload, ncu launch list, ncu --set full of one
# full-size round (3D and 2D; the launch-skip count is read from the launch list), e2e breakdown
cd "$(dirname "$0")/.."
V=${1:-r2b}
O=gpurun_out
PAT="k_attempt_hot|k_commit_coop|k_spheres|k_attempt_slow"
python bench.py > $O/${V}_bench_u3_10m.json 2> $O/${V}_bench_u3_10m.err
for w in u3_1m u2_1m u3_100k u3_10k c3_5m l3_5m b3_64x100k; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/${V}_bench_$w.json 2>&1; done
VOR_STREAM_SETS=1024 python bench.py --workload b3_8192x100k --steps 3 --warmup 1 --no-cpu-baseline > $O/${V}_bench_b3_1024x100k.json 2>&1
python bench.py --workload b3_8192x100k --steps 1 --warmup 1 --no-cpu-baseline > $O/${V}_bench_b3_8192x100k.json 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/${V}_bench_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${V}_launches_u3_10m.csv python tools/one_insert.py 10000000 3 > $O/${V}_ncu_l.log 2>&1
S3=$(python tools/pick_skip.py $O/${V}_launches_u3_10m.csv "$PAT")
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s $S3 -c 4 -o $O/${V}_round_u3_10m -f python tools/one_insert.py 10000000 3 > $O/${V}_ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${V}_launches_u2_4m.csv python tools/one_insert.py 4000000 2 > $O/${V}_ncu_l2.log 2>&1
S2=$(python tools/pick_skip.py $O/${V}_launches_u2_4m.csv "$PAT")
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s $S2 -c 3 -o $O/${V}_round_u2_4m -f python tools/one_insert.py 4000000 2 > $O/${V}_ncu2.log 2>&1
VOR_VERBOSE=1 python tools/e2e_breakdown.py 2>&1 | grep -E "edges:|iter" | tail -6 > $O/${V}_e2e_breakdown.log
echo "skips: 3D $S3 2D $S2"
for f in $O/${V}_bench_*.json; do echo "$f: $(grep '^{' $f | tail -1 | cut -c1-160)"; done
tail -n 2 $O/${V}_ncu3.log; tail -n 2 $O/${V}_ncu2.log

#!/bin/bash
# N-GPU evidence at HEAD (gpurun --gpus N): slab bench and streamed batch over N ranks
cd "$(dirname "$0")/.."
N=${1:-4}
V=r2c
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29621 bench.py --gpus $N --workload u3_10m_slab --steps 4 --warmup 2 > $O/${V}_slab$N.json 2> $O/${V}_slab$N.err
VOR_STREAM_SETS=$((1024*N)) $TR --master-port 29622 bench.py --gpus $N --workload b3_8192x100k --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/${V}_stream$N.json 2> $O/${V}_stream$N.err
for f in slab$N stream$N; do echo "$f: $(grep '^{' $O/${V}_$f.json | tail -1 | cut -c1-200)"; tail -n 2 $O/${V}_$f.err; done
grep '^{' $O/${V}_slab$N.json | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; print(d["n_gpus"], round(d["ms_per_step"],1),"ms", c["tree_points_per_rank"], c["certification_rounds"], c.get("simplices_certified_by_peers"), c["edges_sha256"][:16])'

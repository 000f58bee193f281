#!/bin/bash
cd "$(dirname "$0")/.."
V=r2f
python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -3
for N in 2 1; do
if [ $N = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N"; fi
$L bench.py --gpus $N --workload u3_10m_slab --steps 3 --warmup 2 > gpurun_out/${V}_slab$N.json 2> gpurun_out/${V}_slab$N.err
grep '^{' gpurun_out/${V}_slab$N.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; print(d["n_gpus"], round(d["value"]/1e6,2),"Mpts/s", round(d["ms_per_step"],1),"ms", c["tree_points_per_rank"], c["halo_bytes_received_per_rank"], c["certification_rounds"], c["edges_sha256"][:16])'
tail -2 gpurun_out/${V}_slab$N.err
done

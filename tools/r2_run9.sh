#!/bin/bash
cd "$(dirname "$0")/.."
V=r2e
python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload u3_10m_slab --steps 3 --warmup 1 > gpurun_out/${V}_slab2.json 2> gpurun_out/${V}_slab2.err
tail -c 1800 gpurun_out/${V}_slab2.json; tail -5 gpurun_out/${V}_slab2.err
python bench.py --workload u3_10m_slab --steps 3 --warmup 1 > gpurun_out/${V}_slab1.json 2> gpurun_out/${V}_slab1.err
tail -c 900 gpurun_out/${V}_slab1.json; tail -3 gpurun_out/${V}_slab1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload b3_8192x100k --steps 1 --warmup 1 --no-e2e > gpurun_out/${V}_stream2.json 2> gpurun_out/${V}_stream2.err
tail -c 700 gpurun_out/${V}_stream2.json; tail -3 gpurun_out/${V}_stream2.err

#!/bin/bash
# 2-GPU evidence at HEAD (gpurun --gpus 2): slab tests over NCCL, slab bench 1 / 2 GPUs, streamed batch 2,048 sets on 2 GPUs,
# default workload as 2 replicas, the GPU suite's multi-device test
cd "$(dirname "$0")/.."
V=r2c
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_slab.py tests/test_gpu_engine.py -x -q -k "slab or devices" 2>&1 | tail -3
$TR --master-port 29611 bench.py --gpus 2 --workload u3_10m_slab --steps 3 --warmup 2 > $O/${V}_slab2.json 2> $O/${V}_slab2.err
python bench.py --workload u3_10m_slab --steps 3 --warmup 2 > $O/${V}_slab1.json 2> $O/${V}_slab1.err
VOR_STREAM_SETS=2048 $TR --master-port 29612 bench.py --gpus 2 --workload b3_8192x100k --steps 2 --warmup 1 --no-cpu-baseline > $O/${V}_stream2.json 2> $O/${V}_stream2.err
$TR --master-port 29613 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $O/${V}_main2.json 2> $O/${V}_main2.err
for f in slab2 slab1 stream2 main2; do echo "$f: $(grep '^{' $O/${V}_$f.json | tail -1 | cut -c1-200)"; tail -n 2 $O/${V}_$f.err; done

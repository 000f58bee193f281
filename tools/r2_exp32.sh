#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp32.log; : > $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in 1 2 3; do
echo "## run $i" >> $O
VOR_SLAB_VERBOSE=1 $TR --master-port 2961$i bench.py --gpus 2 --workload u3_10m_slab --steps 1 --warmup 0 2>/dev/null | grep -E "^\[slab|^\{" | cut -c1-260 >> $O
done
cat $O

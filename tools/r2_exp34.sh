#!/bin/bash
# slab mode with the peers' exact in-sphere certificate for the simplices a ball cannot certify
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp34.log; : > $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -2 >> $O
VOR_SLAB_VERBOSE=1 $TR --master-port 29611 bench.py --gpus 2 --workload u3_10m_slab --steps 6 --warmup 2 2>/dev/null > $O.raw
grep -oE "\[slab 1\] round [0-9]+: holds [0-9]+ points[^u]*uncertified [0-9]+ \([0-9]+ after asking the peers\), need \[[^]]*\]" $O.raw | sed -E 's/region .*uncertified/unc/' >> $O
grep '^{' $O.raw | cut -c1-240 >> $O
python bench.py --workload u3_10m_slab --steps 4 --warmup 2 2>/dev/null | grep '^{' | cut -c1-240 >> $O
cat $O

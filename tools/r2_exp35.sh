#!/bin/bash
# what -fmad=false costs: the same sources built with -fmad=true (timing only: parity of the getters' circumsphere needs the unfused form)
cd "$(dirname "$0")/.."
O=gpurun_out/r2_exp35.log; : > $O
for so in voronoids_b200/libvoronoids_b200.so variants/fmad.so voronoids_b200/libvoronoids_b200.so variants/fmad.so; do
echo "## $so" >> $O
VOR_SO=$PWD/$so T_PROFILE=1 python tools/t_stage.py 10000000 3 2>&1 | grep -E "PROFILE|RESULT" >> $O
done
cat $O

#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2_memcheck2.log; : > $O
timeout 60 compute-sanitizer --tool memcheck --print-limit 5 python tools/memcheck_more.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|lattice|batch|incremental|rror" | head -10 >> $O
cat $O

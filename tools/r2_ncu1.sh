#!/bin/bash
cd "$(dirname "$0")/.."
V=${1:-r2b}
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:k_attempt_coop|k_commit_coop|k_spheres" -s 2148 -c 3 -o gpurun_out/${V}_round716 -f python tools/one_insert.py 10000000 3 > gpurun_out/${V}_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/${V}_ncu.log; ls -la gpurun_out/*.ncu-rep

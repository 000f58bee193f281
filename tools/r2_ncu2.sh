#!/bin/bash
cd "$(dirname "$0")/.."
V=${1:-r2c}
export VOR_SO=${2:-voronoids_b200/libvoronoids_b200.so}
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${V}_launches.csv python tools/one_insert.py 10000000 3 > gpurun_out/${V}_ncu_l.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k "regex:k_attempt_hot|k_commit_coop|k_spheres" -s 2148 -c 3 -o gpurun_out/${V}_round716 -f python tools/one_insert.py 10000000 3 > gpurun_out/${V}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${V}_ncu.log
T_PROFILE=1 python tools/t_stage.py 10000000 3 2>&1 | grep -E "PROFILE|RESULT"

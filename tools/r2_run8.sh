#!/bin/bash
cd "$(dirname "$0")/.."
V=r2d
python -m pytest tests -m gpu -x -q > gpurun_out/${V}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${V}_gpu_tests.log
tail -4 gpurun_out/${V}_gpu_tests.log
VOR_STREAM_SETS=1024 python bench.py --workload b3_8192x100k --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${V}_bench_stream1024.json 2> gpurun_out/${V}_bench_stream1024.err
tail -c 1500 gpurun_out/${V}_bench_stream1024.json; tail -3 gpurun_out/${V}_bench_stream1024.err

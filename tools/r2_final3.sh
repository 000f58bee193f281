#!/bin/bash
cd "$(dirname "$0")/.."
python bench.py > gpurun_out/r2d_bench_u3_10m.json 2> gpurun_out/r2d_bench_u3_10m.err
grep '^{' gpurun_out/r2d_bench_u3_10m.json | tail -1 | cut -c1-300

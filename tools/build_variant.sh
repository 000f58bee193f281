#!/bin/bash
# scratch: build a tuning variant of the library: tools/build_variant.sh NAME -DVOR_ATTEMPT_REGS=80 ...  -> variants/NAME.so
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false --expt-relaxed-constexpr \
  -diag-suppress 550 -diag-suppress 63 -Xcompiler -fPIC -Xcompiler -pthread -shared "$@" -Iinclude \
  -o variants/$name.so voronoids_b200/csrc/vor_lib.cu > variants/$name.log 2>&1
echo "$name rc=$?"

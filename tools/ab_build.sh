#!/bin/bash
# Development: A/B builds of libb200align.so with compile-time switches -> build/ab/<name>.so (select with B200_LIB).
#   tools/ab_build.sh name "-DFLAG ..."
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $2 -shared -o build/ab/$1.so masa-cudalign_b200/csrc/engine.cu

#!/bin/bash
# Builds A/B variants of the library into variants/ (git-ignored) for tools/ab_bench.py.
set -e
cd "$(dirname "$0")/../se-195-project-ray-tracer_b200"
rm -f ../variants/*.so
build() { tag=$1; shift; nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -std=c++17 -Xcompiler -fPIC,-O2,-fno-fast-math,-ffp-contract=off -shared -o ../variants/librt_$tag.so csrc/rt_kernels.cu csrc/rt_api.cu host/scene_io.cpp & }
build base
build w128b5 -DW_MIN_BLOCKS=5
build w128b6 -DW_MIN_BLOCKS=6
build w64b10 -DW_THREADS=64 -DW_MIN_BLOCKS=10
build w256b2 -DW_THREADS=256 -DW_MIN_BLOCKS=2
build w64 -DW_THREADS=64
wait
ls ../variants

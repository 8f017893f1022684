#!/bin/bash
# Builds A/B variants of the library into variants/ (git-ignored) for tools/ab_bench.py.
set -e
cd "$(dirname "$0")/../se-195-project-ray-tracer_b200"
build() { tag=$1; shift; nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -std=c++17 -Xcompiler -fPIC,-O2,-fno-fast-math,-ffp-contract=off -shared -o ../variants/librt_$tag.so csrc/rt_kernels.cu csrc/rt_api.cu host/scene_io.cpp & }
build base
build w128b8 -DW_MIN_BLOCKS=8
build w128b6 -DW_MIN_BLOCKS=6
build w256b3 -DW_THREADS=256 -DW_MIN_BLOCKS=3
build w64 -DW_THREADS=64
build nopair -DW_PLANE_PAIRS=0
build pt128 -DPT_THREADS=128
build pt512 -DPT_THREADS=512
build pt256b5 -DPT_MIN_BLOCKS=5
wait
ls -la ../variants

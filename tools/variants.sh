#!/bin/bash
# Builds A/B variants of the library into variants/ (git-ignored) for tools/ab_bench.py.
set -e
cd "$(dirname "$0")/../se-195-project-ray-tracer_b200"
rm -f ../variants/*.so
build() { tag=$1; shift; nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -std=c++17 -Xcompiler -fPIC,-O2,-fno-fast-math,-ffp-contract=off -shared -o ../variants/librt_$tag.so csrc/rt_kernels.cu csrc/rt_api.cu host/scene_io.cpp & }
build na -DW_PLANE_PAIRS=0 -DW_NO_AXIS
build na_b11 -DW_PLANE_PAIRS=0 -DW_NO_AXIS -DW_MIN_BLOCKS=11
build na_b10 -DW_PLANE_PAIRS=0 -DW_NO_AXIS -DW_MIN_BLOCKS=10
build ax_b11 -DW_PLANE_PAIRS=0 -DW_MIN_BLOCKS=11
build ax_b10 -DW_PLANE_PAIRS=0 -DW_MIN_BLOCKS=10
build na2_b11 -DW_PLANE_PAIRS=0 -DW_SPHERE_PAIRS=0 -DW_NO_AXIS -DW_MIN_BLOCKS=11
wait
ls ../variants

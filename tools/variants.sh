#!/bin/bash
# Builds A/B variants of the library into variants/ (git-ignored) for tools/ab_bench.py.
set -e
cd "$(dirname "$0")/../se-195-project-ray-tracer_b200"
rm -f ../variants/*.so
build() { tag=$1; shift; nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -std=c++17 -Xcompiler -fPIC,-O2,-fno-fast-math,-ffp-contract=off -shared -o ../variants/librt_$tag.so csrc/rt_kernels.cu csrc/rt_api.cu host/scene_io.cpp & }
build prod
build stats -DW_ROUND_STATS
# examples: build b10 -DW_MIN_BLOCKS=10 ; build leaf4 -DPT_BVH_LEAF_MAX=4 ; build nomargin -DPT_BVH_TEST_NO_MARGIN (tools/bvh_fuzz.py self-check)
wait
ls ../variants

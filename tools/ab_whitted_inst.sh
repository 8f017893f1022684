#!/bin/bash
# A/B of library variants (variants/librt_*.so + the product) on the 1080p Whitted frame: CUDA-event time, then the
# executed-instruction count and lanes per instruction of whitted_kernel from ncu (numbers under ncu are not timings).
cd "$(dirname "$0")/.."
for so in "$@"; do
  export RT_B200_LIB=$so
  [ "$so" = product ] && unset RT_B200_LIB
  python tools/ab_bench.py
  ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum --clock-control none -k regex:whitted_kernel -c 2 --csv python tools/prof_run.py whitted 2>/dev/null | grep whitted_kernel | awk -F'","' '{print "   ncu:", $(NF-2), $(NF)}' | tr -d '"'
done

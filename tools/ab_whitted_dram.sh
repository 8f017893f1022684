#!/bin/bash
# DRAM traffic, instructions and time of whitted_kernel (1080p) per library variant, from ncu (experiments; not timings)
cd "$(dirname "$0")/.."
for so in "$@"; do
  echo "== $so"
  if [ "$so" = product ]; then unset RT_B200_LIB; else export RT_B200_LIB=$so; fi
  ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_write.sum --clock-control none -k regex:whitted_kernel -s 1 -c 1 --csv python tools/prof_run.py whitted 2>/dev/null | grep whitted_kernel | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $(NF)}' | tr -d '"'
done

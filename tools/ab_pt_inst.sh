#!/bin/bash
# executed instructions / lanes / time of pt_kernel (Cornell 1024x768 x 16 spp) per library variant, from ncu (not a timing)
cd "$(dirname "$0")/.."
for so in "$@"; do
  echo "== $so"
  if [ "$so" = product ]; then unset RT_B200_LIB; else export RT_B200_LIB=$so; fi
  ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_active --clock-control none -k regex:pt_kernel -c 2 --csv python tools/prof_run.py pt 2>/dev/null | grep pt_kernel | awk -F'","' '{print "   ", $(NF-2), $(NF)}' | tr -d '"'
done

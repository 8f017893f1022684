#!/bin/bash
# compute-sanitizer over tools/sanitize_run.py: memcheck, racecheck, synccheck, initcheck; summaries into gpurun_out/r02_sanitizer_<tool>.txt
cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck initcheck; do
  out=gpurun_out/r02_sanitizer_$tool.txt
  echo "== compute-sanitizer --tool $tool python tools/sanitize_run.py" > $out
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 $( [ $tool = memcheck ] && echo "--leak-check full" ) python tools/sanitize_run.py 2>&1 | grep -vE "^\s*$" | tail -40 >> $out
  echo "exit: ${PIPESTATUS[0]}" >> $out
done
out=gpurun_out/r02_sanitizer_memcheck_ipc.txt
echo "== compute-sanitizer --tool memcheck --target-processes all python tools/sanitize_run.py ipc" > $out
timeout 900 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 python tools/sanitize_run.py ipc 2>&1 | grep -vE "^\s*$" | tail -40 >> $out
echo "exit: ${PIPESTATUS[0]}" >> $out
tail -n 6 gpurun_out/r02_sanitizer_*.txt

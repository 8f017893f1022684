#!/bin/bash
# Device-side bounds checks of our own (the GPU pool refuses compute-sanitizer): builds the library with -DRT_DEVICE_CHECKS and, as
# a self-check, once more with a FIFO limit of 4 slots that the glass pixels MUST exceed; then runs tools/sanitize_run.py against both.
#   tools/checked_build.sh build   (here, no GPU)      tools/checked_build.sh run   (on the GPU box; prints the verdicts)
cd "$(dirname "$0")/.."
case "$1" in
build)
  mkdir -p variants
  make -C se-195-project-ray-tracer_b200 librt_b200.so OUT=../variants/librt_checked.so EXTRA=-DRT_DEVICE_CHECKS -B > /dev/null 2>&1 &
  make -C se-195-project-ray-tracer_b200 librt_b200.so OUT=../variants/librt_checked_selfcheck.so EXTRA="-DRT_DEVICE_CHECKS -DW_QUEUE_CHECK_SLOTS=4" -B > /dev/null 2>&1 &
  wait; ls -la variants/librt_checked*.so ;;
run)
  echo "== product build (checks not compiled in)"; python tools/sanitize_run.py 2>&1 | tail -1
  echo "== -DRT_DEVICE_CHECKS"; RT_B200_LIB=variants/librt_checked.so python tools/sanitize_run.py 2>&1 | tail -1
  echo "== -DRT_DEVICE_CHECKS, two ranks, rank 1 stores into rank 0's frame through CUDA IPC"; RT_B200_LIB=$PWD/variants/librt_checked.so python tools/sanitize_run.py ipc 2>&1 | grep -E "rank|ok"
  echo "== self-check: -DRT_DEVICE_CHECKS -DW_QUEUE_CHECK_SLOTS=4 (the FIFO check, bit 0, must fire)"; RT_B200_LIB=variants/librt_checked_selfcheck.so python tools/sanitize_run.py 2>&1 | tail -1 ;;
*) echo "usage: $0 build|run" ;;
esac

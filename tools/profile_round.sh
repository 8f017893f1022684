#!/bin/bash
# The round's measurement set on one B200 box (run through gpurun): GPU tests, smoke(), the bench line of both arms, then -- each only
# after the same command has exited 0 without ncu -- the ncu launch list of the bench and `ncu --set full` captures of the three workloads
# of tools/prof_run.py.  Everything lands in gpurun_out/; tools/ncu_summary.py / ncu_funcs.py condense the captures afterwards.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -z "$PROFILE_ONLY" ] || [ "$PROFILE_ONLY" = none ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gputests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02f_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r02f_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference_arm.log 2>&1; echo "reference arm rc=$?"
python bench.py > gpurun_out/r02f_bench_1gpu.log 2>gpurun_out/r02f_bench_1gpu.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/r02f_bench_steps5.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_ncu_launches_bench_steps5.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r02f_ncu_launches.log 2>&1
fi
[ "$PROFILE_ONLY" = none ] && exit 0
# (gpurun brings back at most 64 MiB per call: run the captures as  PROFILE_ONLY="whitted" / "pt bvh4"  in calls of their own)
for wl in ${PROFILE_ONLY:-whitted pt bvh4}; do
  k=whitted; [ $wl = pt ] && k=pt_kernel; [ $wl = bvh4 ] && k=pt_bvh_kernel
  python tools/prof_run.py $wl > /dev/null 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:$k -c ${PROFILE_LAUNCHES:-12} -f -o gpurun_out/r02f_$wl python tools/prof_run.py $wl > gpurun_out/r02f_ncu_$wl.log 2>&1
done
ls -la gpurun_out | tail -20

#!/bin/bash
# ncu evidence for the bench command (run under gpurun; see /opt/skills/guides/B200_PROFILING.md).
#   1. plain run (must exit 0)  2. launch list with per-launch device time
#   3. full capture of the ten K3 launches of one 256-scan step: the 128-thread variant of the kernel is only
#      used by the batched steps; bench order with --no-latency: correctness step (10 launches), warm-up step
#      (the next 10) ...
#   TAG=v8 bash profiles/run_ncu.sh            (c2, the bench default)
#   TAG=v8 WL=c1 bash profiles/run_ncu.sh      (REFERENCE mode: full capture of ref_search / ref_reduce / ref_step)
set -e
TAG=${TAG:-v8}
WL=${WL:-c2}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency --workload $WL"
[ -n "$SPS" ] && CMD="$CMD --scans-per-step $SPS"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}_$WL.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}_$WL.csv $CMD > gpurun_out/ncu_launches_${TAG}_$WL.log 2>&1
if [ "$WL" = "c2" ]; then
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:search_accum_kernel<.*128>' -s 10 -c 10 -f -o gpurun_out/search_accum_$TAG $CMD > gpurun_out/ncu_full_${TAG}_$WL.log 2>&1
else
  ncu --set full --clock-control none --import-source on -k 'regex:ref_(search|reduce|step)_kernel' -s 12 -c 12 -f -o gpurun_out/ref_kernels_$TAG $CMD > gpurun_out/ncu_full_${TAG}_$WL.log 2>&1
fi
tail -1 gpurun_out/plain_${TAG}_$WL.log

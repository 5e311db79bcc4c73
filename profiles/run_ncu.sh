#!/bin/bash
# ncu evidence for the bench command (run under gpurun; see /opt/skills/guides/B200_PROFILING.md).
#   1. plain run (must exit 0)  2. launch list with per-launch device time
#   3. full capture of the ten K3 launches of one 256-scan step (bench order: 10 launches of the
#      correctness step, 5 x 10 of the single-scan latency probe, then the warm-up step = launches 60..69)
set -e
TAG=${TAG:-v6}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --scans-per-step ${SPS:-256}"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_c2.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_accum -s 60 -c 10 -f -o gpurun_out/search_accum_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log

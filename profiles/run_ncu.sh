#!/bin/bash
# ncu evidence for the bench command (run under gpurun; see /opt/skills/guides/B200_PROFILING.md).
#   1. plain run (must exit 0)  2. launch list with per-launch device time  3. full capture of K3
set -e
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --scans-per-step ${SPS:-64}"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_accum -s 12 -c 2 -f -o gpurun_out/prof_search $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/plain.log

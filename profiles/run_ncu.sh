#!/bin/bash
# Profiling recipe (B200_PROFILING.md).  Run under gpurun from the repo root; outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
# 1. launch list with per-launch device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
# 2. full capture of the dominant kernel (one launch, small batch keeps the ~40 replays short)
ncu --set full --clock-control none --import-source on -k regex:step_random -s 1 -c 1 -f -o gpurun_out/${TAG}_step_random \
    python bench.py --steps 1 --warmup 1 --games 16384 --no-cpu-baseline > gpurun_out/${TAG}_full_bench.log 2>&1
ncu -i gpurun_out/${TAG}_step_random.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_random_raw.csv 2>/dev/null
ls -la gpurun_out

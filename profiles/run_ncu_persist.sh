#!/bin/bash
mkdir -p gpurun_out
RV_WARPS_PER_SM=8 RV_ACT_REPS=4 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rollout_persistent -s 1 -c 1 -f -o gpurun_out/r01_persist \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01_persist_bench.log 2>&1
ncu -i gpurun_out/r01_persist.ncu-rep --page raw --csv > gpurun_out/r01_persist_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_persist.ncu-rep --page source --csv --print-source sass > gpurun_out/r01_persist_source.csv 2>/dev/null
ls -la gpurun_out | tail -4

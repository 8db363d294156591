#!/bin/bash
# Round-1 final profiling pass (run under gpurun from the repo root).
mkdir -p gpurun_out
# 1. launch list of the steady state of one rollout (phase pipeline): 400 consecutive launches
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400 --csv --log-file gpurun_out/r01_final_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r01_final_launches_bench.log 2>&1
# 2. full capture: one launch of the ACT fast kernel and one of the generic (SLOW) kernel in the steady state
ncu --set full --clock-control none --import-source on -k regex:phase_kernel -s 1200 -c 8 -f -o gpurun_out/r01_final_phase \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r01_final_full_bench.log 2>&1
ncu -i gpurun_out/r01_final_phase.ncu-rep --page raw --csv > gpurun_out/r01_final_phase_raw.csv 2>/dev/null
ls -la gpurun_out | tail -6

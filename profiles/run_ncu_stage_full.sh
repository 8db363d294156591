#!/bin/bash
mkdir -p gpurun_out
RV_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase_kernel -s 1500 -c 6 -f -o gpurun_out/r01_stage_phase \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r01_stage_full_bench.log 2>&1
ncu -i gpurun_out/r01_stage_phase.ncu-rep --page raw --csv > gpurun_out/r01_stage_phase_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_stage_phase.ncu-rep --page source --csv --print-source sass > gpurun_out/r01_stage_phase_source.csv 2>/dev/null
ls -la gpurun_out | tail -4

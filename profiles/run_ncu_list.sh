#!/bin/bash
# launch list of the steady state of one rollout (phase pipeline, graph off so every launch is a plain kernel)
mkdir -p gpurun_out
RV_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 360 --csv --log-file gpurun_out/r01_stage_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r01_stage_launches_bench.log 2>&1
tail -2 gpurun_out/r01_stage_launches.csv | cut -c1-200
RV_DEBUG=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-300

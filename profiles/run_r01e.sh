#!/bin/bash
mkdir -p gpurun_out
for wps in 10 6; do
RV_WARPS_PER_SM=$wps RV_MONO_BELOW=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3500 -c 700 --csv --log-file gpurun_out/r01e_launches_$wps.csv \
    python bench.py --workload rollout_obs --unfused --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01e_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01e_launches_$wps.csv | head -12
done
RV_MONO_BELOW=0 timeout 300 python bench.py --workload rollout_obs --unfused --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01e_bench.json 2> gpurun_out/r01e.err
cut -c1-250 gpurun_out/r01e_bench.json; tail -3 gpurun_out/r01e.err

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01l_bench_obs.json 2>> gpurun_out/r01l.err
cut -c1-250 gpurun_out/r01l_bench_obs.json
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 2500 -c 300 --csv --log-file gpurun_out/r01l_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01l_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01l_obs_launches.csv | head -8
grep "step_sorted_kernel" gpurun_out/r01l_obs_launches.csv | grep inst_executed | head -2 | cut -c1-40,200-400

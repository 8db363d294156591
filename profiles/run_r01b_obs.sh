#!/bin/bash
# Session-2 pass: gpu tests at HEAD, main bench, and a launch list of the rollout_obs workload (mid-game launches).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err
cut -c1-300 gpurun_out/r01b_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/r01b_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_obs_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01b_obs_launches.csv | head -20

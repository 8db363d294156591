#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload rollout_obs --split --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_bench_obs_split.json 2> gpurun_out/r01d.err
cut -c1-250 gpurun_out/r01d_bench_obs_split.json; tail -3 gpurun_out/r01d.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 500 --csv --log-file gpurun_out/r01d_obs_split_launches.csv \
    python bench.py --workload rollout_obs --split --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01d_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01d_obs_split_launches.csv | head -12
timeout 300 python bench.py --workload hands --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_bench_hands.json 2>> gpurun_out/r01d.err
cut -c1-200 gpurun_out/r01d_bench_hands.json; grep -o '"e2e": {[^}]*}' gpurun_out/r01d_bench_hands.json
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hand" 2>&1 | tail -3

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "observ" 2>&1 | tail -5
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01j_bench_obs.json 2>> gpurun_out/r01j.err
cut -c1-250 gpurun_out/r01j_bench_obs.json
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 2500 -c 300 --csv --log-file gpurun_out/r01j_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01j_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01j_obs_launches.csv | head -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_random_kernel -s 300 -c 1 -f -o gpurun_out/r01j_step \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01j_ncu.log 2>&1
ncu -i gpurun_out/r01j_step.ncu-rep --page raw --csv > gpurun_out/r01j_step_raw.csv 2>/dev/null
ncu -i gpurun_out/r01j_step.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r01j_step_source.csv 2>/dev/null
ls -la gpurun_out | grep r01j

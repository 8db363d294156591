#!/bin/bash
# Fused observe+step kernel: parity, bench (fused and unfused), launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "observ" 2>&1 | tail -15
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_bench_obs.json 2> gpurun_out/r01c_bench_obs.err
cut -c1-250 gpurun_out/r01c_bench_obs.json; tail -3 gpurun_out/r01c_bench_obs.err
timeout 300 python bench.py --workload rollout_obs --unfused --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_bench_obs_unfused.json 2>> gpurun_out/r01c_bench_obs.err
cut -c1-250 gpurun_out/r01c_bench_obs_unfused.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/r01c_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01c_obs_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01c_obs_launches.csv | head -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/r01c_obs_unfused_launches.csv \
    python bench.py --workload rollout_obs --unfused --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01c_obs_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01c_obs_unfused_launches.csv | head -12

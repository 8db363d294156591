#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "observ" 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01h_bench.json 2> gpurun_out/r01h.err
cut -c1-200 gpurun_out/r01h_bench.json; tail -3 gpurun_out/r01h.err
timeout 300 python bench.py --mode 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01h_bench_3p.json 2>> gpurun_out/r01h.err
cut -c1-200 gpurun_out/r01h_bench_3p.json
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01h_bench_obs.json 2>> gpurun_out/r01h.err
cut -c1-250 gpurun_out/r01h_bench_obs.json
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 2500 -c 300 --csv --log-file gpurun_out/r01h_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01h_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01h_obs_launches.csv | head -8
grep "obs_encode_kernel" gpurun_out/r01h_obs_launches.csv | grep inst_executed | head -3

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "observ" 2>&1 | tail -5
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_bench_obs.json 2> gpurun_out/r01f.err
cut -c1-250 gpurun_out/r01f_bench_obs.json; tail -3 gpurun_out/r01f.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 600 --csv --log-file gpurun_out/r01f_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01f_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r01f_obs_launches.csv | head -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obs_encode_kernel -s 300 -c 1 -f -o gpurun_out/r01f_encode \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01f_ncu.log 2>&1
ncu -i gpurun_out/r01f_encode.ncu-rep --page raw --csv > gpurun_out/r01f_encode_raw.csv 2>/dev/null
ncu -i gpurun_out/r01f_encode.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r01f_encode_source.csv 2>/dev/null
ls -la gpurun_out | grep r01f

#!/bin/bash
# Full capture of the dominant kernel only (see run_ncu.sh for the launch list).
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:step_random -s 1 -c 1 -f -o gpurun_out/${TAG}_step_random \
    python bench.py --steps 1 --warmup 1 --games ${2:-65536} --no-cpu-baseline > gpurun_out/${TAG}_full_bench.log 2>&1
ncu -i gpurun_out/${TAG}_step_random.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_random_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_step_random.ncu-rep --page source --csv > gpurun_out/${TAG}_step_random_source.csv 2>/dev/null
ls -la gpurun_out | tail -5

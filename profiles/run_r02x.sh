#!/bin/bash
# round 2, session x: smoke(), and what bounds the one-step kernel of the observation pipeline (full ncu capture mid-rollout)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02x_smoke.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 600 --csv --log-file gpurun_out/r02x_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02x_obs_launches_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02x_obs_launches.csv > gpurun_out/r02x_obs_launches_summary.txt; head -9 gpurun_out/r02x_obs_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_staged -s 300 -c 1 -f -o gpurun_out/r02x_staged \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02x_staged_bench.log 2>&1
ncu -i gpurun_out/r02x_staged.ncu-rep --page raw --csv > gpurun_out/r02x_staged_raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/r02x_staged_raw.csv 0 > gpurun_out/r02x_staged_ncu_summary.txt 2>&1; head -40 gpurun_out/r02x_staged_ncu_summary.txt

#!/bin/bash
# round 2, session c: profiles before the kernel work of this round — hand_eval_kernel (never captured in round 1) and the
# persistent rollout kernel (baseline), ncu --set full with sources
mkdir -p gpurun_out
python bench.py --workload hands --steps 3 --warmup 1 > gpurun_out/r02c_bench_hands.json 2> gpurun_out/r02c_bench_hands.err
ncu --set full --clock-control none --import-source on -k regex:hand_eval_kernel -s 1 -c 1 -f -o gpurun_out/r02c_hands \
    python bench.py --workload hands --steps 1 --warmup 1 > gpurun_out/r02c_hands_ncu.log 2>&1
ncu -i gpurun_out/r02c_hands.ncu-rep --page raw --csv > gpurun_out/r02c_hands_raw.csv 2>/dev/null
ncu -i gpurun_out/r02c_hands.ncu-rep --page source --csv > gpurun_out/r02c_hands_source.csv 2>/dev/null
cat gpurun_out/r02c_bench_hands.json
ls -la gpurun_out | tail -6
# config 5 after the warp-cooperative deal: launch list of the observe+step pipeline
python bench.py --workload rollout_obs --steps 2 --warmup 1 > gpurun_out/r02c_bench_rollout_obs.json 2> gpurun_out/r02c_bench_rollout_obs.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02c_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 0 --games 65536 > gpurun_out/r02c_obs_launches_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02c_obs_launches.csv > gpurun_out/r02c_obs_launches_summary.txt 2>&1
cat gpurun_out/r02c_bench_rollout_obs.json; cat gpurun_out/r02c_obs_launches_summary.txt | head -20
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "observe or multi_device" 2>&1 | tail -3

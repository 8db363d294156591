#!/bin/bash
# round 2, session d: two-kernel hand evaluation, staged one-step kernel (config 5), multi-device handle
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hand_eval or lockstep or observe or multi_device or external or partial" 2>&1 | tail -5 > gpurun_out/r02d_pytest.txt
cat gpurun_out/r02d_pytest.txt
python bench.py --workload hands --steps 3 --warmup 1 > gpurun_out/r02d_bench_hands.json 2> gpurun_out/r02d_bench_hands.err
cat gpurun_out/r02d_bench_hands.json
for m in 2 3; do
RV_STEP_SORTED=$m python bench.py --workload rollout_obs --steps 2 --warmup 1 > gpurun_out/r02d_bench_rollout_obs_sorted$m.json 2> gpurun_out/r02d_bench_rollout_obs_sorted$m.err
cat gpurun_out/r02d_bench_rollout_obs_sorted$m.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1200 --csv --log-file gpurun_out/r02d_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 0 --games 65536 > gpurun_out/r02d_obs_launches_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02d_obs_launches.csv > gpurun_out/r02d_obs_launches_summary.txt 2>&1
head -12 gpurun_out/r02d_obs_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:hand_ -s 2 -c 2 -f -o gpurun_out/r02d_hands \
    python bench.py --workload hands --steps 1 --warmup 1 > gpurun_out/r02d_hands_ncu.log 2>&1
ncu -i gpurun_out/r02d_hands.ncu-rep --page raw --csv > gpurun_out/r02d_hands_raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/r02d_hands_raw.csv > gpurun_out/r02d_hands_ncu_summary.txt 2>&1
head -60 gpurun_out/r02d_hands_ncu_summary.txt

#!/bin/bash
# round 2, session h: A/B of the headline rollout (this build vs the build of session a), hands after the head-walk change,
# watchdog test, sanitizers
mkdir -p gpurun_out
python profiles/ab_rollout.py tmp_r02a.so riichienv_b200/libriichienv_b200.so > gpurun_out/r02h_ab_rollout.txt 2>&1
cat gpurun_out/r02h_ab_rollout.txt
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "watchdog or hand_eval" 2>&1 | tail -3 > gpurun_out/r02h_pytest.txt
cat gpurun_out/r02h_pytest.txt
python bench.py --workload hands --steps 3 --warmup 1 > gpurun_out/r02h_bench_hands.json 2> gpurun_out/r02h_bench_hands.err
cut -c1-200 gpurun_out/r02h_bench_hands.json
ncu --set full --clock-control none --import-source on -k regex:hand_yaku -s 1 -c 1 -f -o gpurun_out/r02h_yaku \
    python bench.py --workload hands --steps 1 --warmup 1 > gpurun_out/r02h_yaku_ncu.log 2>&1
ncu -i gpurun_out/r02h_yaku.ncu-rep --page raw --csv > gpurun_out/r02h_yaku_raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/r02h_yaku_raw.csv 0 > gpurun_out/r02h_hands_yaku_ncu_summary.txt 2>&1
head -24 gpurun_out/r02h_hands_yaku_ncu_summary.txt
bash profiles/run_r02_sanitize.sh > /dev/null 2>&1
cat gpurun_out/r02_sanitize.txt | cut -c1-200

#!/bin/bash
# round 2, session i: rollout scheduler with one parallel control read per visit + claim loops over set bits.
# A/B against the build of the previous commit in ONE process, parity gates, per-class accounting, full ncu capture with source.
mkdir -p gpurun_out
python profiles/ab_rollout.py tmp_head.so riichienv_b200/libriichienv_b200.so > gpurun_out/r02i_ab_rollout.txt 2>&1
cat gpurun_out/r02i_ab_rollout.txt
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_games or parity_gate or greedy or settlement or watchdog or partial or lockstep or sanma_config" 2>&1 | tail -3 > gpurun_out/r02i_pytest.txt
cat gpurun_out/r02i_pytest.txt
RV_LIB_PATH=$PWD/tmp_qprof.so timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02i_qprof_bench.json 2> gpurun_out/r02i_qprof.err
grep qprof gpurun_out/r02i_qprof.err | tail -18 > gpurun_out/r02i_qprof.txt
cat gpurun_out/r02i_qprof.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_persistent -s 1 -c 1 -f -o gpurun_out/r02i_persist \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02i_persist_bench.log 2>&1
ncu -i gpurun_out/r02i_persist.ncu-rep --page raw --csv > gpurun_out/r02i_persist_raw.csv 2>/dev/null
ncu -i gpurun_out/r02i_persist.ncu-rep --page source --csv --print-source sass > gpurun_out/r02i_persist_source.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/r02i_persist_raw.csv 0 > gpurun_out/r02i_persist_ncu_summary.txt 2>&1
head -40 gpurun_out/r02i_persist_ncu_summary.txt

#!/bin/bash
# round 2, session m: crew kernel vs per-warp scheduler with the knobs fixed (0 is a value now), ncu of the crew kernel
mkdir -p gpurun_out
L=riichienv_b200/libriichienv_b200.so
python profiles/ab_rollout.py $L:RV_CREW=0 $L:RV_CREW=1 $L:RV_CREW=1,RV_INIT_DIST=0 $L:RV_CREW=0,RV_ACT_HOLD=1 $L:RV_CREW=1,RV_WARPS_PER_SM=12 > gpurun_out/r02m_ab_rollout.txt 2>&1
cat gpurun_out/r02m_ab_rollout.txt
RV_WARPS_PER_SM=8 python profiles/ab_rollout.py $L:RV_CREW=1 $L:RV_CREW=0 > gpurun_out/r02m_ab_rollout_w8.txt 2>&1
cat gpurun_out/r02m_ab_rollout_w8.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_crew -s 1 -c 1 -f -o gpurun_out/r02m_crew \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02m_crew_bench.log 2>&1
ncu -i gpurun_out/r02m_crew.ncu-rep --page raw --csv > gpurun_out/r02m_crew_raw.csv 2>/dev/null
ncu -i gpurun_out/r02m_crew.ncu-rep --page source --csv --print-source sass > gpurun_out/r02m_crew_source.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/r02m_crew_raw.csv 0 > gpurun_out/r02m_crew_ncu_summary.txt 2>&1
head -40 gpurun_out/r02m_crew_ncu_summary.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
cut -c1-300 gpurun_out/r02m_bench.json

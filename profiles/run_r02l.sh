#!/bin/bash
# round 2, session l: the crew kernel (one block per SM, warps in lock-step) against the per-warp scheduler, same build
mkdir -p gpurun_out
L=riichienv_b200/libriichienv_b200.so
python profiles/ab_rollout.py $L:RV_CREW=0,RV_ACT_HOLD=0 $L:RV_CREW=1 $L:RV_CREW=1,RV_ACT_REPS=2 $L:RV_CREW=1,RV_ACT_REPS=8 $L:RV_CREW=1,RV_SWITCH_IDLE=10 \
   $L:RV_CREW=1,RV_INIT_DIST=1 > gpurun_out/r02l_ab_rollout.txt 2>&1
cat gpurun_out/r02l_ab_rollout.txt
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_games or watchdog or partial or greedy_agent" 2>&1 | tail -3 > gpurun_out/r02l_pytest.txt
cat gpurun_out/r02l_pytest.txt

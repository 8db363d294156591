#!/bin/bash
# round 2, session t: plain turns taken at the end of TAIL / DEAL / SLOW visits (follow reps) in the crew kernel
mkdir -p gpurun_out
L=riichienv_b200/libriichienv_b200.so
python profiles/ab_rollout.py $L $L:RV_FOLLOW_REPS=1 $L:RV_FOLLOW_REPS=2 $L:RV_FOLLOW_REPS=4 $L:RV_FOLLOW_REPS=2,RV_ACT_REPS=6 $L:RV_FOLLOW_REPS=4,RV_ACT_REPS=8 2>&1 | tail -8 | cut -c1-200 | tee gpurun_out/r02t_ab_rollout.txt
RV_FOLLOW_REPS=2 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_games or watchdog or partial or greedy_agent or encode_100k" 2>&1 | tail -3 | tee gpurun_out/r02t_pytest.txt

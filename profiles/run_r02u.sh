#!/bin/bash
# round 2, session u: visits of several rounds (plain turns / discard follow-ups alternating under block barriers)
mkdir -p gpurun_out
L=riichienv_b200/libriichienv_b200.so
python profiles/ab_rollout.py $L $L:RV_ROUNDS=1 $L:RV_ROUNDS=2 $L:RV_ROUNDS=3 $L:RV_ROUNDS=4 $L:RV_ROUNDS=6 $L:RV_ROUNDS=3,RV_PHASE_SYNC=0 $L:RV_ROUNDS=3,RV_ACT_REPS=2 $L:RV_ROUNDS=3,RV_ACT_REPS=8 2>&1 | tail -10 | cut -c1-200 | tee gpurun_out/r02u_ab_rollout.txt
RV_ROUNDS=3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_games or watchdog or partial or greedy_agent" 2>&1 | tail -3 | tee gpurun_out/r02u_pytest.txt

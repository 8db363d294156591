#!/bin/bash
# round 2, session j: lane refill in the persistent rollout kernel (ACT -> ACT games stay staged in their lane).
# A/B in one process: build of the previous session, this build with the refill off / on and 1, 2, 4, 8 act_fast reps per iteration
mkdir -p gpurun_out
L=riichienv_b200/libriichienv_b200.so
python profiles/ab_rollout.py tmp_head.so $L:RV_ACT_HOLD=0 $L:RV_ACT_HOLD=1 $L:RV_ACT_HOLD=1,RV_ACT_REPS=2 $L:RV_ACT_HOLD=1,RV_ACT_REPS=1 \
    $L:RV_ACT_HOLD=1,RV_ACT_REPS=8 > gpurun_out/r02j_ab_rollout.txt 2>&1
cat gpurun_out/r02j_ab_rollout.txt
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_games or watchdog or partial or greedy_agent" 2>&1 | tail -3 > gpurun_out/r02j_pytest.txt
cat gpurun_out/r02j_pytest.txt
python -m pytest tests/test_replay.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r02j_pytest_replay.txt
cat gpurun_out/r02j_pytest_replay.txt

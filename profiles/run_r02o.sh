#!/bin/bash
# round 2, session o: crew kernel with the decision taken one iteration ahead (prefetch) and the greedier class choice
mkdir -p gpurun_out
L=riichienv_b200/libriichienv_b200.so
python profiles/ab_rollout.py $L:RV_CREW_PREFETCH=0,RV_CREW_GREEDY=0 $L:RV_CREW_PREFETCH=1,RV_CREW_GREEDY=0 $L:RV_CREW_PREFETCH=0,RV_CREW_GREEDY=1 $L \
   $L:RV_ACT_REPS=3 $L:RV_ACT_REPS=6 $L:RV_ENDGAME_Q=4 $L:RV_ENDGAME_Q=16 > gpurun_out/r02o_ab_rollout.txt 2>&1
cat gpurun_out/r02o_ab_rollout.txt
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_games or watchdog or partial or greedy_agent" 2>&1 | tail -3 > gpurun_out/r02o_pytest.txt
cat gpurun_out/r02o_pytest.txt

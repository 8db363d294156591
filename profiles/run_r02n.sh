#!/bin/bash
# round 2, session n: where the crew kernel's warps wait (barrier accounting) and a greedier class choice
mkdir -p gpurun_out
L=riichienv_b200/libriichienv_b200.so
python profiles/ab_rollout.py $L $L:RV_CREW_GREEDY=1 $L:RV_SWITCH_IDLE=1 $L:RV_CREW_GREEDY=1,RV_SWITCH_IDLE=1 > gpurun_out/r02n_ab_rollout.txt 2>&1
cat gpurun_out/r02n_ab_rollout.txt
for knobs in "RV_CREW=1" "RV_CREW=1 RV_CREW_GREEDY=1"; do
  echo "== $knobs" >> gpurun_out/r02n_qprof.txt
  env $knobs RV_LIB_PATH=$PWD/tmp_qprof.so timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep qprof | tail -19 >> gpurun_out/r02n_qprof.txt
done
cat gpurun_out/r02n_qprof.txt

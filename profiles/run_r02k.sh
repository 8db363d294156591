#!/bin/bash
# round 2, session k: how often an SM changes class (every change refills its instruction cache).
# Sweep of the switch rule (empty looks before a change, minimum length of the target queue), the initial distribution, lane refill.
mkdir -p gpurun_out
L=riichienv_b200/libriichienv_b200.so
python profiles/ab_rollout.py $L:RV_ACT_HOLD=0 $L:RV_ACT_HOLD=0,RV_SWITCH_IDLE=10 $L:RV_ACT_HOLD=0,RV_SWITCH_IDLE=30 $L:RV_ACT_HOLD=0,RV_SWITCH_IDLE=100 \
   $L:RV_ACT_HOLD=0,RV_SWITCH_IDLE=30,RV_SWITCH_MINLEN=64 $L:RV_ACT_HOLD=0,RV_SWITCH_IDLE=30,RV_SWITCH_MINLEN=512 \
   $L:RV_ACT_HOLD=0,RV_SWITCH_IDLE=30,RV_SWITCH_MINLEN=64,RV_INIT_DIST=1 $L:RV_ACT_HOLD=1,RV_SWITCH_IDLE=30,RV_SWITCH_MINLEN=64,RV_INIT_DIST=1 \
   $L:RV_ACT_HOLD=0,RV_SWITCH_IDLE=300,RV_SWITCH_MINLEN=64,RV_INIT_DIST=1 \
   > gpurun_out/r02k_ab_rollout.txt 2>&1
cat gpurun_out/r02k_ab_rollout.txt
for knobs in "RV_SWITCH_IDLE=3" "RV_SWITCH_IDLE=30 RV_SWITCH_MINLEN=64 RV_INIT_DIST=1" "RV_SWITCH_IDLE=300 RV_SWITCH_MINLEN=64 RV_INIT_DIST=1"; do
  echo "== $knobs" >> gpurun_out/r02k_qprof.txt
  env $knobs RV_ACT_HOLD=0 RV_LIB_PATH=$PWD/tmp_qprof.so timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep qprof | tail -18 >> gpurun_out/r02k_qprof.txt
done
cat gpurun_out/r02k_qprof.txt

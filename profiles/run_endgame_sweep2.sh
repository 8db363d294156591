#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1; }
run RV_ENDGAME_Q=1 RV_ENDGAME_TAKE=1
run RV_ENDGAME_Q=2 RV_ENDGAME_TAKE=1
run RV_ENDGAME_Q=4 RV_ENDGAME_TAKE=1
run RV_ENDGAME_Q=6 RV_ENDGAME_TAKE=1
run RV_ENDGAME_Q=8 RV_ENDGAME_TAKE=1
run RV_ENDGAME_Q=4 RV_ENDGAME_TAKE=2
run RV_ENDGAME_Q=4 RV_ENDGAME_TAKE=1 RV_WARPS_PER_SM=12
run RV_ENDGAME_Q=4 RV_ENDGAME_TAKE=1 RV_ACT_REPS=8
run RV_ENDGAME_Q=4 RV_ENDGAME_TAKE=1 RV_ACT_REPS=2
echo "== 3P"; timeout 300 python bench.py --mode 5 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1
RV_LIB_PATH=$PWD/tmp_qprof.so timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/qprof2_bench.json 2> gpurun_out/qprof2.err
grep qprof gpurun_out/qprof2.err | tail -17

#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1; }
run RV_ENDGAME_OFF=1
run RV_ENDGAME_Q=4 RV_ENDGAME_TAKE=4
run RV_ENDGAME_Q=8 RV_ENDGAME_TAKE=4
run RV_ENDGAME_Q=16 RV_ENDGAME_TAKE=4
run RV_ENDGAME_Q=16 RV_ENDGAME_TAKE=8
run RV_ENDGAME_Q=32 RV_ENDGAME_TAKE=8
run RV_ENDGAME_Q=64 RV_ENDGAME_TAKE=16
run RV_ENDGAME_Q=128 RV_ENDGAME_TAKE=32
run RV_ENDGAME_Q=8 RV_ENDGAME_TAKE=2
run RV_ENDGAME_Q=16 RV_ENDGAME_TAKE=2
run RV_ENDGAME_Q=4 RV_ENDGAME_TAKE=1
run RV_ENDGAME_OFF=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rollout or hanchan or results or gate" 2>&1 | tail -3

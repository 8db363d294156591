#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1; }
run RV_POST_REPS=1
run RV_POST_REPS=2
run RV_POST_REPS=3
run RV_POST_REPS=5
run RV_POST_REPS=3 RV_ACT_REPS=6
run RV_POST_REPS=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rollout or hanchan or results or gate" 2>&1 | tail -3

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "random_games or partial_rollout or lockstep or observe_step" 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1; done
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in 4 6 8 10; do for r in 2 4; do
  echo "== WARPS_PER_SM=$w ACT_REPS=$r"; RV_WARPS_PER_SM=$w RV_ACT_REPS=$r timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -oE "\"value\": [0-9.]+|Error.*|error.*" | head -2
done; done

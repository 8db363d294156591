#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in 11 12 12 11; do
  echo "WARPS_PER_SM=$w: $(RV_WARPS_PER_SM=$w timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1)"
done
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1
timeout 300 python bench.py --mode 5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for g in 1 0; do
  echo "== RV_GRAPH=$g"; RV_DEBUG=1 RV_GRAPH=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "iterations|\"value\"" | tail -2 | cut -c1-400
done
for se in 2 4; do for dm in 4 8 16; do
  echo "== SLOW_EVERY=$se DEAL_MULT=$dm"; RV_SLOW_EVERY=$se RV_DEAL_MULT=$dm timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -oE "\"value\": [0-9.]+" | head -1
done; done

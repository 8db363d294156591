#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for r in 1 2 4 8; do
  echo "== RV_ACT_REPS=$r"; RV_DEBUG=1 RV_ACT_REPS=$r timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -oE "iterations=[0-9]+|\"value\": [0-9.]+" | head -3 | tr '\n' ' '; echo
done
for se in 1 2; do for dm in 4 8 16; do
  echo "== SLOW_EVERY=$se DEAL_MULT=$dm"; RV_SLOW_EVERY=$se RV_DEAL_MULT=$dm timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -oE "\"value\": [0-9.]+" | head -1
done; done

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
b() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -oE "\"value\": [0-9.]+|iterations=[0-9]+" | tail -2 | tr '\n' ' '; echo; }
for se in 1 2; do for r in 2 4 8; do
  echo "== phased SLOW_EVERY=$se REPS=$r"; RV_DEBUG=1 RV_ROLLOUT=phased RV_SLOW_EVERY=$se RV_ACT_REPS=$r b
done; done
for w in 3 4 6 10; do for r in 4 8; do
  echo "== persistent W=$w REPS=$r"; RV_WARPS_PER_SM=$w RV_ACT_REPS=$r b
done; done
echo "== mono"; RV_ROLLOUT=mono b

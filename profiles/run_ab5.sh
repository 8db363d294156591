#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
b() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -oE "\"value\": [0-9.]+|rror.*" | head -1; }
for w in 6 8 10; do for r in 3 4 6; do
  echo "== persistent W=$w REPS=$r $(RV_WARPS_PER_SM=$w RV_ACT_REPS=$r b)"
done; done
for w in 4 10; do echo "== qprof W=$w"; RV_LIB_PATH=$PWD/tmp_qprof.so RV_WARPS_PER_SM=$w RV_ACT_REPS=8 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "qprof" | tail -17 | cut -c1-200; done

#!/bin/bash
for kv in "RV_ENDGAME_Q=4" "RV_ENDGAME_Q=8" "RV_ENDGAME_Q=12" "RV_ACT_REPS=3" "RV_ACT_REPS=6" "RV_ENDGAME_TAKE=2" "RV_ENDGAME_Q=8"; do
  echo "$kv: $(env $kv timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1)"
done

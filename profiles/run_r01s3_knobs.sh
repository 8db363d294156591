#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r01s3
RV_LIB_PATH=$PWD/tmp_qprof.so timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_qprof_bench.json 2> ${O}_qprof.err
grep qprof ${O}_qprof.err | tail -18 | tee ${O}_qprof.txt
for reps in 2 4 8 16; do
  echo "ACT_REPS=$reps: $(RV_ACT_REPS=$reps timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1)"
done
for q in 4 8 16; do
  echo "ENDGAME_Q=$q: $(RV_ENDGAME_Q=$q timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+' | head -1)"
done

#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_games_vs_oracle or sequence or lockstep" 2>&1 | tail -2
for i in 1 2; do timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -oE "\"value\": [0-9.]+" | head -1; done

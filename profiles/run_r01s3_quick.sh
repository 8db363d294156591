#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+|"ms_per_step": [0-9.]+' | head -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -oE '"value": [0-9.]+|"ms_per_step": [0-9.]+' | head -2

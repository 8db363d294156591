#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400

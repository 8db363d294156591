#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r01s3h_pytest_gpu.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r01s3h_bench.json 2>/dev/null; cut -c1-200 gpurun_out/r01s3h_bench.json

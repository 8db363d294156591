#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --mode 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01g_bench_3p.json 2> gpurun_out/r01g.err
cut -c1-200 gpurun_out/r01g_bench_3p.json; tail -3 gpurun_out/r01g.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01g_bench.json 2>> gpurun_out/r01g.err
cut -c1-200 gpurun_out/r01g_bench.json

#!/bin/bash
# round 2, session a: GPU test suite (incl. the reference's own pytest suite against the product) + bench regression check
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest_gpu.txt
python tests/refsuite/run.py gpu > gpurun_out/r02a_refsuite_gpu.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
cat gpurun_out/r02a_pytest_gpu.txt; tail -3 gpurun_out/r02a_refsuite_gpu.txt; cat gpurun_out/r02a_bench.json

#!/bin/bash
# round 2, session b: full GPU suite (greedy-agent gates, MJAI text parity on all viewers, reference suite on the product)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02b_pytest_gpu.txt
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "settlement_gate" 2>&1 | grep "greedy gate" > gpurun_out/r02b_greedy_hist.txt
python tests/refsuite/run.py gpu > gpurun_out/r02b_refsuite_gpu.txt 2>&1
cat gpurun_out/r02b_pytest_gpu.txt; cat gpurun_out/r02b_greedy_hist.txt; tail -2 gpurun_out/r02b_refsuite_gpu.txt

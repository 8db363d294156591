#!/bin/bash
# round 2, session r: full GPU suite on the crew-kernel build (incl. the reference's own suite through the shim and the replay tests)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02r_pytest_gpu.txt
cat gpurun_out/r02r_pytest_gpu.txt

#!/bin/bash
mkdir -p gpurun_out
RV_LIB_PATH=$PWD/tmp_qprof.so timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/qprof_bench.json 2> gpurun_out/qprof.err
cut -c1-200 gpurun_out/qprof_bench.json
grep qprof gpurun_out/qprof.err | tail -20

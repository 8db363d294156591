#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r01s3
timeout 120 python profiles/sanitize_small.py 256 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file ${O}_memcheck.txt python profiles/sanitize_small.py 128 > ${O}_memcheck_run.txt 2>&1
echo "memcheck rc=$?"; tail -3 ${O}_memcheck_run.txt; tail -5 ${O}_memcheck.txt

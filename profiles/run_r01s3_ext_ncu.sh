#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r01s3
timeout 300 python -m pytest tests -m gpu -q -k "shim_observation_encode_extended" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:obs_ext_kernel -s 2 -c 1 -f -o ${O}_ext \
    python profiles/time_encode_ext.py 65536 > ${O}_ext_ncu.log 2>&1
ncu -i ${O}_ext.ncu-rep --page raw --csv > ${O}_ext_raw.csv 2>/dev/null
ncu -i ${O}_ext.ncu-rep --page source --csv > ${O}_ext_source.csv 2>/dev/null
ls -la ${O}_ext*

#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r01s3
timeout 900 python -m pytest tests -m gpu -q -k "extended or ext_ or observation_encode or observe_step" 2>&1 | tail -25 | tee ${O}_pytest_ext.txt
timeout 300 python profiles/time_encode_ext.py > ${O}_time_encode_ext.json 2> ${O}_time.err; cat ${O}_time_encode_ext.json; tail -3 ${O}_time.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file ${O}_ext_launches.csv \
    python profiles/time_encode_ext.py 65536 > /dev/null 2>&1
python profiles/summarize_launches.py ${O}_ext_launches.csv > ${O}_ext_launches_summary.txt; head -7 ${O}_ext_launches_summary.txt
grep "obs_ext_kernel" ${O}_ext_launches.csv | grep inst_executed | head -1 | awk -F'","' '{print "ext inst", $NF}'

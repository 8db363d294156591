#!/bin/bash
# Round-1 (session 3) final measurement pass; run under gpurun from the repo root; writes gpurun_out/r01s2_*.
mkdir -p gpurun_out
O=gpurun_out/r01s3g
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee ${O}_pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
tail -c 1800 ${O}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference.json 2>> ${O}_bench.err
cut -c1-400 ${O}_bench_reference.json
timeout 300 python bench.py --mode 5 --steps 3 --warmup 3 --no-cpu-baseline > ${O}_bench_3p.json 2>> ${O}_bench.err
timeout 300 python bench.py --workload hands --steps 3 --warmup 3 --no-cpu-baseline > ${O}_bench_hands.json 2>> ${O}_bench.err
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline > ${O}_bench_rollout_obs.json 2>> ${O}_bench.err
timeout 300 python bench.py --workload rollout_obs --mode 5 --steps 2 --warmup 3 --no-cpu-baseline > ${O}_bench_rollout_obs_3p.json 2>> ${O}_bench.err
for f in 3p hands rollout_obs rollout_obs_3p; do grep -oE "\"metric\": \"[a-z_]+\", \"value\": [0-9.]+" ${O}_bench_$f.json; done
# launch list of the default bench command (1 warm-up + 2 timed rollouts)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_launches_bench.log 2>&1
python profiles/summarize_launches.py ${O}_launches.csv > ${O}_launches_summary.txt; head -8 ${O}_launches_summary.txt
# full capture of one rollout launch
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rollout_persistent -s 1 -c 1 -f -o ${O}_persist \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > ${O}_persist_bench.log 2>&1
ncu -i ${O}_persist.ncu-rep --page raw --csv > ${O}_persist_raw.csv 2>/dev/null
# observation pipeline: launch list (mid-game) and full capture of the encoder
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 600 --csv --log-file ${O}_obs_launches.csv \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > ${O}_obs_launches_bench.log 2>&1
python profiles/summarize_launches.py ${O}_obs_launches.csv > ${O}_obs_launches_summary.txt; head -9 ${O}_obs_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:obs_encode_kernel -s 300 -c 1 -f -o ${O}_encode \
    python bench.py --workload rollout_obs --steps 1 --warmup 1 --no-cpu-baseline > ${O}_encode_bench.log 2>&1
ncu -i ${O}_encode.ncu-rep --page raw --csv > ${O}_encode_raw.csv 2>/dev/null
ls -la gpurun_out | grep r01s2
# extended encoder: timing + launch list
timeout 300 python profiles/time_encode_ext.py > ${O}_time_encode_ext.json 2>> ${O}_bench.err; cat ${O}_time_encode_ext.json
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file ${O}_ext_launches.csv \
    python profiles/time_encode_ext.py 65536 > /dev/null 2>&1
python profiles/summarize_launches.py ${O}_ext_launches.csv > ${O}_ext_launches_summary.txt; head -6 ${O}_ext_launches_summary.txt
ls gpurun_out | grep r01s3f | wc -l

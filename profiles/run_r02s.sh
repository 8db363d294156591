#!/bin/bash
# round 2, session s: measurement pass on the crew-kernel build — bench lines of every workload, the reference arm, the ncu
# launch list and full capture of the dominant kernel, sanitizers.  Run under gpurun from the repo root.
mkdir -p gpurun_out
O=gpurun_out/r02s
timeout 600 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
tail -c 2200 ${O}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference.json 2>> ${O}_bench.err
cut -c1-400 ${O}_bench_reference.json
timeout 300 python bench.py --mode 5 --steps 3 --warmup 3 --no-cpu-baseline > ${O}_bench_3p.json 2>> ${O}_bench.err
timeout 300 python bench.py --workload hands --steps 3 --warmup 3 --no-cpu-baseline > ${O}_bench_hands.json 2>> ${O}_bench.err
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline > ${O}_bench_rollout_obs.json 2>> ${O}_bench.err
timeout 300 python bench.py --workload rollout_obs --mode 5 --steps 2 --warmup 3 --no-cpu-baseline > ${O}_bench_rollout_obs_3p.json 2>> ${O}_bench.err
for f in 3p hands rollout_obs rollout_obs_3p; do grep -oE "\"metric\": \"[a-z_]+\", \"value\": [0-9.]+" ${O}_bench_$f.json; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_launches_bench.log 2>&1
python profiles/summarize_launches.py ${O}_launches.csv > ${O}_launches_summary.txt; head -8 ${O}_launches_summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rollout_crew -s 1 -c 1 -f -o ${O}_crew \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > ${O}_crew_bench.log 2>&1
ncu -i ${O}_crew.ncu-rep --page raw --csv > ${O}_crew_raw.csv 2>/dev/null
ncu -i ${O}_crew.ncu-rep --page source --csv --print-source sass > ${O}_crew_source.csv 2>/dev/null
python profiles/summarize_ncu.py ${O}_crew_raw.csv 0 > ${O}_crew_ncu_summary.txt 2>&1; head -36 ${O}_crew_ncu_summary.txt
rm -f gpurun_out/r02_sanitize.txt
bash profiles/run_r02_sanitize.sh > /dev/null 2>&1
cut -c1-160 gpurun_out/r02_sanitize.txt

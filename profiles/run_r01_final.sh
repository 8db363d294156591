#!/bin/bash
# Round-1 final measurement pass (run under gpurun from the repo root; writes gpurun_out/r01f_*).
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err
tail -c 1500 gpurun_out/r01f_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01f_bench_reference.json 2>> gpurun_out/r01f_bench.err
cat gpurun_out/r01f_bench_reference.json | cut -c1-600
timeout 300 python bench.py --mode 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_bench_3p.json 2>> gpurun_out/r01f_bench.err
timeout 300 python bench.py --workload hands --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_bench_hands.json 2>> gpurun_out/r01f_bench.err
timeout 300 python bench.py --workload rollout_obs --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_bench_obs.json 2>> gpurun_out/r01f_bench.err
for f in 3p hands obs; do grep -oE "\"metric\": \"[a-z_]+\", \"value\": [0-9.]+" gpurun_out/r01f_bench_$f.json; done
# launch list of the default bench command (every kernel launch of 1 warm-up + 2 timed rollouts)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01f_launches_bench.log 2>&1
# full capture of one rollout launch
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rollout_persistent -s 1 -c 1 -f -o gpurun_out/r01f_persist \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01f_persist_bench.log 2>&1
ncu -i gpurun_out/r01f_persist.ncu-rep --page raw --csv > gpurun_out/r01f_persist_raw.csv 2>/dev/null
ncu -i gpurun_out/r01f_persist.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r01f_persist_source.csv 2>/dev/null
ls -la gpurun_out | grep r01f

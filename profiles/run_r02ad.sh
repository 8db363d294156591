#!/bin/bash
# round 2, final: headline bench + reference arm + launch list + full capture of the crew kernel on the final sources
mkdir -p gpurun_out
O=gpurun_out/r02ad
timeout 600 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
tail -c 1500 ${O}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference.json 2>> ${O}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_launches_bench.log 2>&1
python profiles/summarize_launches.py ${O}_launches.csv > ${O}_launches_summary.txt; head -6 ${O}_launches_summary.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rollout_crew -s 1 -c 1 -f -o ${O}_crew \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > ${O}_crew_bench.log 2>&1
ncu -i ${O}_crew.ncu-rep --page raw --csv > ${O}_crew_raw.csv 2>/dev/null
rm -f ${O}_crew.ncu-rep        # (gpurun brings back at most 64 MiB)
python profiles/summarize_ncu.py ${O}_crew_raw.csv 0 > ${O}_crew_ncu_summary.txt 2>&1; head -24 ${O}_crew_ncu_summary.txt
python -m pytest tests/test_gpu_parity.py tests/test_replay.py tests/test_gpu_shim.py -m gpu -q -k "random_games or watchdog or partial or replay or shim or multi_device" 2>&1 | tail -2 | tee ${O}_pytest.txt
timeout 300 python bench.py --workload hands --steps 3 --warmup 3 --no-cpu-baseline > ${O}_bench_hands.json 2>> ${O}_bench.err
grep -oE "\"metric\": \"[a-z_]+\", \"value\": [0-9.]+" ${O}_bench_hands.json
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "hand_eval or shanten" 2>&1 | tail -2 | tee ${O}_pytest_hands.txt
timeout 600 ncu --set full --clock-control none -k regex:hand_yaku -s 1 -c 1 -f -o ${O}_yaku python bench.py --workload hands --steps 1 --warmup 1 --no-cpu-baseline > ${O}_yaku_ncu.log 2>&1
ncu -i ${O}_yaku.ncu-rep --page raw --csv > ${O}_yaku_raw.csv 2>/dev/null
rm -f ${O}_yaku.ncu-rep
python profiles/summarize_ncu.py ${O}_yaku_raw.csv 0 > ${O}_hands_yaku_ncu_summary.txt 2>&1; head -36 ${O}_hands_yaku_ncu_summary.txt | tail -22

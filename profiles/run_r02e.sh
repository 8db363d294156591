#!/bin/bash
# round 2, session e: hand evaluation after the candidate-list change (ncu of both kernels), GPU parity of hands
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hand_eval or greedy_agent" 2>&1 | tail -3 > gpurun_out/r02e_pytest.txt
cat gpurun_out/r02e_pytest.txt
python bench.py --workload hands --steps 3 --warmup 1 > gpurun_out/r02e_bench_hands.json 2> gpurun_out/r02e_bench_hands.err
cat gpurun_out/r02e_bench_hands.json
ncu --set full --clock-control none --import-source on -k regex:hand_ -s 2 -c 2 -f -o gpurun_out/r02e_hands \
    python bench.py --workload hands --steps 1 --warmup 1 > gpurun_out/r02e_hands_ncu.log 2>&1
ncu -i gpurun_out/r02e_hands.ncu-rep --page raw --csv > gpurun_out/r02e_hands_raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/r02e_hands_raw.csv 1 > gpurun_out/r02e_hands_yaku_ncu_summary.txt 2>&1
python profiles/summarize_ncu.py gpurun_out/r02e_hands_raw.csv 0 > gpurun_out/r02e_hands_shape_ncu_summary.txt 2>&1
head -40 gpurun_out/r02e_hands_yaku_ncu_summary.txt

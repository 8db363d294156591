#!/bin/bash
# round 2, session f: observation pipeline with the encoder overlapped with the step kernel (A/B), parity, hands after the DFS change,
# ncu --set full of the persistent rollout kernel of this build (-> profiles/ncu_facts.json)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shim.py -m gpu -q -x -k "observe or encode or obs or shim or hand_eval_golden" 2>&1 | tail -3 > gpurun_out/r02f_pytest.txt
cat gpurun_out/r02f_pytest.txt
for m in 1 2; do
RV_OBS_OVERLAP=$m python bench.py --workload rollout_obs --steps 2 --warmup 1 > gpurun_out/r02f_bench_rollout_obs_overlap$m.json 2> gpurun_out/r02f_bench_rollout_obs_overlap$m.err
cut -c1-220 gpurun_out/r02f_bench_rollout_obs_overlap$m.json
done
python bench.py --workload hands --steps 3 --warmup 1 > gpurun_out/r02f_bench_hands.json 2> gpurun_out/r02f_bench_hands.err
cut -c1-220 gpurun_out/r02f_bench_hands.json
ncu --set full --clock-control none --import-source on -k regex:rollout_persistent -s 1 -c 1 -f -o gpurun_out/r02f_persist \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02f_persist_bench.log 2>&1
ncu -i gpurun_out/r02f_persist.ncu-rep --page raw --csv > gpurun_out/r02f_persist_raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/r02f_persist_raw.csv 0 > gpurun_out/r02f_persist_ncu_summary.txt 2>&1
head -45 gpurun_out/r02f_persist_ncu_summary.txt

#!/bin/bash
# round 2, session g: full GPU suite (apply_event on the product, overlap, staged kernel), refsuite on the GPU, config 5 A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02g_pytest_gpu.txt
cat gpurun_out/r02g_pytest_gpu.txt
python tests/refsuite/run.py gpu > gpurun_out/r02g_refsuite_gpu.txt 2>&1
tail -1 gpurun_out/r02g_refsuite_gpu.txt
for m in 1 2; do
RV_OBS_OVERLAP=$m python bench.py --workload rollout_obs --steps 2 --warmup 1 > gpurun_out/r02g_bench_rollout_obs_overlap$m.json 2> gpurun_out/r02g_bench_rollout_obs_overlap$m.err
cut -c1-200 gpurun_out/r02g_bench_rollout_obs_overlap$m.json; tail -2 gpurun_out/r02g_bench_rollout_obs_overlap$m.err
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
cut -c1-200 gpurun_out/r02g_bench.json

#!/bin/bash
# round 2, session w: several GPUs of one box (run as: gpurun --gpus N -- bash profiles/run_r02w_multi.sh N)
# headline at N ranks, BASELINE configs[4] (tensors every step, 125,000 games per GPU = 10^6 games on 8 GPUs), and the in-library
# multi-device handle (results must not depend on the device count)
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r02w_n$N
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
cut -c1-330 ${O}_bench.json
timeout 900 $TR bench.py --workload rollout_obs --gpus $N --games 125000 --steps 2 --warmup 1 > ${O}_bench_rollout_obs.json 2> ${O}_bench_rollout_obs.err
cut -c1-330 ${O}_bench_rollout_obs.json; tail -2 ${O}_bench_rollout_obs.err
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device" 2>&1 | tail -3 > ${O}_pytest_multi.txt
cat ${O}_pytest_multi.txt
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > ${O}_gpus.txt

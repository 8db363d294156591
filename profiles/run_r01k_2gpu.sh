#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r01k_bench_2gpu.json 2> gpurun_out/r01k.err
tail -c 2500 gpurun_out/r01k_bench_2gpu.json; tail -5 gpurun_out/r01k.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r01k_bench_2gpu_ref.json 2>> gpurun_out/r01k.err
tail -c 600 gpurun_out/r01k_bench_2gpu_ref.json

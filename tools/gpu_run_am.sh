#!/bin/bash
# round-2 session am (2 GPUs): BASELINE config 4 (wave RK4, 256^3 cells per GPU, Jacobi-CG mass solves), weak scaling point N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload wave_rk4 --steps 5 --warmup 1 > gpurun_out/am_wave_2gpu.json 2> gpurun_out/am_wave_2gpu.err

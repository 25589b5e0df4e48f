#!/bin/bash
# round-2 session ad: BASELINE configs 3 and 4 at their full per-GPU sizes on one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python bench.py --workload wave_rk4 --steps 10 > gpurun_out/ad_wave256.json 2> gpurun_out/ad_wave256.err
timeout 400 python bench.py --workload advection_rk4 --steps 5 --warmup 1 > gpurun_out/ad_adv512.json 2> gpurun_out/ad_adv512.err

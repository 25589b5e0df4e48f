#!/bin/bash
# round-2 session ah: launch list of one RK4 wave step with the direct mass inverse (per-kernel times)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/ah_launches_wave_direct.csv python bench.py --workload wave_rk4 --steps 3 --warmup 1 --mass-solver direct > gpurun_out/ah_ncu.log 2>&1

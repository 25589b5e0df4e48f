#!/bin/bash
# round-2 session ai: batched loads in the line solves -- parity, RK workloads with the direct mass inverse
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_solvers.py -q -x -k "mass_inverse" > gpurun_out/ai_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ai_pytest.log
timeout 100 python bench.py --workload wave_rk4 --steps 10 --mass-solver direct > gpurun_out/ai_wave256_direct.json 2> gpurun_out/ai_wave256_direct.err
timeout 200 python bench.py --workload advection_rk4 --steps 5 --warmup 1 --mass-solver direct > gpurun_out/ai_adv512_direct.json 2> gpurun_out/ai_adv512_direct.err

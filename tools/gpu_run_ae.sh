#!/bin/bash
# round-2 session ae: Kronecker-direct mass inverse -- parity tests, RK workloads (BASELINE configs 3, 4) with it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_solvers.py -q -x -k "mass_inverse" > gpurun_out/ae_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ae_pytest.log
timeout 200 python bench.py --workload wave_rk4 --steps 10 --mass-solver direct > gpurun_out/ae_wave256_direct.json 2> gpurun_out/ae_wave256_direct.err
timeout 300 python bench.py --workload advection_rk4 --steps 5 --warmup 1 --mass-solver direct > gpurun_out/ae_adv512_direct.json 2> gpurun_out/ae_adv512_direct.err

#!/bin/bash
# round-2 session al: smoke + quick bench of the library rebuilt from a clean tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/al_smoke.log 2>&1
timeout 60 python bench.py --quick --steps 200 --warmup 10 > gpurun_out/al_bench.json 2> gpurun_out/al_bench.err
timeout 60 python -m pytest tests/test_gpu_solvers.py -q -x -k "mass_inverse or poisson_01" > gpurun_out/al_pytest.log 2>&1

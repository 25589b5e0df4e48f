#!/bin/bash
# round-2 session ag: last sanity pass of the final tree (solver / apply tests, smoke, quick bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 260 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_apply.py -q -x > gpurun_out/ag_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ag_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ag_smoke.log 2>&1
timeout 60 python bench.py --quick --steps 300 --warmup 20 > gpurun_out/ag_bench.json 2> gpurun_out/ag_bench.err

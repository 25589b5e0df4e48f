#!/bin/bash
# round-2 session ao (last GPU seconds of the round): the cut Poisson tests on the GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 24 python -m pytest tests/test_gpu_cut.py -q -x -s > gpurun_out/ao_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/ao_pytest.log

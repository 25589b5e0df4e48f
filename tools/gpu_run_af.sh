#!/bin/bash
# round-2 session af: mass-inverse parity after the short-block fix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_solvers.py -q -k "mass_inverse" > gpurun_out/af_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/af_pytest.log

#!/bin/bash
# v4 kernel: parity tests, then throughput of v3 vs v4 variants on the BASELINE size
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --deselect "tests/test_gpu_solvers.py::test_cg_iteration_parity_with_oracle" > gpurun_out/pytest_v4.log 2>&1
tail -5 gpurun_out/pytest_v4.log
timeout 200 python -m pytest tests/test_gpu_solvers.py -m gpu -q -k test_cg_iteration_parity_with_oracle > gpurun_out/pytest_cgpar.log 2>&1
grep -E "assert|last_step|passed|failed" gpurun_out/pytest_cgpar.log | head -20
run() { r=$(env "$@" GDM_FUSED_VERBOSE=1 timeout 120 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "gdm\]|value|rror" | tr '\n' ' '); echo "$* :: $r"; }
{
run GDM_FUSED_FAMILY=3
run GDM_FUSED_CFG=100
run GDM_FUSED_CFG=100 GDM_FUSED_RSPLIT=0
run GDM_FUSED_CFG=103
run GDM_FUSED_CFG=104
run GDM_FUSED_CFG=105
run GDM_FUSED_CFG=106
run GDM_FUSED_CFG=108
run GDM_FUSED_CFG=110
} > gpurun_out/v4_tune1.log 2>&1
cat gpurun_out/v4_tune1.log

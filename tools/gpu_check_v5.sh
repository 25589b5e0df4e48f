#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" GDM_FUSED_VERBOSE=1 timeout 120 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "gdm\]|value|rror" | tr '\n' ' ' | cut -c1-400); echo "$* :: $r"; }
{
run GDM_FUSED_CFG=200
run GDM_FUSED_CFG=200 GDM_FUSED_DBG=1
run GDM_FUSED_CFG=200 GDM_FUSED_DBG=4
run GDM_FUSED_CFG=200 GDM_FUSED_DBG=8
run GDM_FUSED_CFG=200 GDM_FUSED_DBG=12
run GDM_FUSED_CFG=203
run GDM_FUSED_CFG=204
run GDM_FUSED_CFG=206
run GDM_FUSED_CFG=207
run GDM_FUSED_CFG=200 GDM_FUSED_SLOTS=148
run GDM_FUSED_CFG=100
} > gpurun_out/v5_tune1.log 2>&1
cat gpurun_out/v5_tune1.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect "tests/test_gpu_solvers.py::test_cg_iteration_parity_with_oracle" > gpurun_out/pytest_v5.log 2>&1
tail -5 gpurun_out/pytest_v5.log

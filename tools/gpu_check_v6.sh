#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" GDM_FUSED_VERBOSE=1 timeout 120 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "gdm\]|value|rror" | tr '\n' ' ' | cut -c1-330); echo "$* :: $r"; }
{
run GDM_FUSED_CFG=300
run GDM_FUSED_CFG=300 GDM_FUSED_DBG=1
run GDM_FUSED_CFG=300 GDM_FUSED_DBG=8
run GDM_FUSED_CFG=302
run GDM_FUSED_CFG=303
run GDM_FUSED_CFG=304
run GDM_FUSED_CFG=305
run GDM_FUSED_CFG=306
run GDM_FUSED_CFG=307
run GDM_FUSED_CFG=309
run GDM_FUSED_CFG=310
run GDM_FUSED_CFG=300 GDM_FUSED_RSPLIT=0
} > gpurun_out/v6_tune1.log 2>&1
cat gpurun_out/v6_tune1.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect "tests/test_gpu_solvers.py::test_cg_iteration_parity_with_oracle" > gpurun_out/pytest_v6.log 2>&1
tail -5 gpurun_out/pytest_v6.log

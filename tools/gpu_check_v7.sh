#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" GDM_FUSED_VERBOSE=1 timeout 40 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "gdm\]|value|rror" | tr '\n' ' ' | sed -E 's/"unit.*//; s/\{"metric": "gdm_stiffness_apply_3d_p3_fp64", //' | cut -c1-260); echo "$* :: $r"; }
{
run GDM_FUSED_CFG=400
r1=$(tail -1 gpurun_out/v7_tune2.log)
if echo "$r1" | grep -q value; then
run GDM_FUSED_CFG=404
run GDM_FUSED_CFG=407
run GDM_FUSED_CFG=408
run GDM_FUSED_CFG=409
run GDM_FUSED_CFG=410
run GDM_FUSED_CFG=411
run GDM_FUSED_CFG=400 GDM_FUSED_SLOTS=256
fi
} > gpurun_out/v7_tune2.log 2>&1
cat gpurun_out/v7_tune2.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 --deselect "tests/test_gpu_solvers.py::test_cg_iteration_parity_with_oracle" > gpurun_out/pytest_v7.log 2>&1
tail -3 gpurun_out/pytest_v7.log

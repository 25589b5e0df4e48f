#!/bin/bash
mkdir -p gpurun_out
GDM_FUSED_CFG=142 timeout 200 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout 120 -k "test_fused_apply_matches_oracle and 3-" > gpurun_out/pytest_ys.log 2>&1; tail -2 gpurun_out/pytest_ys.log
run() { r=$(env "$@" GDM_FUSED_VERBOSE=1 timeout 40 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "gdm\]|value|rror" | tr '\n' ' ' | sed -E 's/"unit.*//; s/\{"metric": "gdm_stiffness_apply_3d_p3_fp64", //' | cut -c1-200); echo "$* :: $r"; }
{
run GDM_FUSED_CFG=142
run GDM_FUSED_CFG=141
run GDM_FUSED_CFG=140
} > gpurun_out/v4_yscatter.log 2>&1
cat gpurun_out/v4_yscatter.log

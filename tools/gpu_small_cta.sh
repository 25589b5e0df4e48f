#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" GDM_FUSED_VERBOSE=1 timeout 60 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "gdm\]|value|rror" | tr '\n' ' ' | sed -E 's/"unit.*//; s/\{"metric": "gdm_stiffness_apply_3d_p3_fp64", //' | cut -c1-200); echo "$* :: $r"; }
{
for c in 130 131 132 133 134; do run GDM_FUSED_CFG=$c; done
run GDM_FUSED_CFG=130 GDM_FUSED_DBG=32
} > gpurun_out/v4_small_cta.log 2>&1
cat gpurun_out/v4_small_cta.log

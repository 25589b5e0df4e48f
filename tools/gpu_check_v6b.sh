#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" GDM_FUSED_VERBOSE=1 timeout 40 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "gdm\]|value|rror" | tr '\n' ' ' | sed -E 's/"unit.*//' | cut -c1-200); echo "$* :: $r"; }
{
run GDM_FUSED_CFG=304
r1=$(tail -1 gpurun_out/v6_tune2.log)
if echo "$r1" | grep -q value; then
run GDM_FUSED_CFG=304 GDM_FUSED_DBG=16
run GDM_FUSED_CFG=311 GDM_FUSED_DBG=16
run GDM_FUSED_CFG=312 GDM_FUSED_DBG=16
run GDM_FUSED_CFG=312 GDM_FUSED_DBG=16 GDM_FUSED_LZ=64
run GDM_FUSED_CFG=312 GDM_FUSED_DBG=16 GDM_FUSED_LZ=52
run GDM_FUSED_CFG=313 GDM_FUSED_DBG=16
run GDM_FUSED_CFG=314 GDM_FUSED_DBG=16
run GDM_FUSED_CFG=315 GDM_FUSED_DBG=16
run GDM_FUSED_CFG=316 GDM_FUSED_DBG=16
run GDM_FUSED_CFG=312
fi
} > gpurun_out/v6_tune2.log 2>&1
cat gpurun_out/v6_tune2.log

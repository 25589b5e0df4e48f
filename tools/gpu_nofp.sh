#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" timeout 40 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "value|rror" | python -c "import sys,json; [print(round(json.loads(l)['value'],1), round(json.loads(l)['ms_per_step'],4)) for l in sys.stdin]"); echo "$* :: $r"; }
{
run GDM_FUSED_CFG=100 GDM_FUSED_LZ=29 GDM_FUSED_DBG=0
run GDM_FUSED_CFG=100 GDM_FUSED_LZ=29 GDM_FUSED_DBG=32
run GDM_FUSED_CFG=100 GDM_FUSED_LZ=29 GDM_FUSED_DBG=33
run GDM_FUSED_CFG=100 GDM_FUSED_LZ=29 GDM_FUSED_DBG=36
run GDM_FUSED_CFG=100 GDM_FUSED_LZ=29 GDM_FUSED_DBG=40
run GDM_FUSED_CFG=100 GDM_FUSED_LZ=29 GDM_FUSED_DBG=12
} > gpurun_out/v4_nofp.log 2>&1
cat gpurun_out/v4_nofp.log

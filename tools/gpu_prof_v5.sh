#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" timeout 120 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "value|rror" | python -c "import sys,json; [print(round(json.loads(l)['value'],1), round(json.loads(l)['ms_per_step'],4)) for l in sys.stdin]"); echo "$* :: $r"; }
{
for dbg in 0 4 8 12; do
run GDM_FUSED_CFG=200 GDM_FUSED_DBG=$dbg
done
} > gpurun_out/v5_ablate.log 2>&1
cat gpurun_out/v5_ablate.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:kron3d -s 10 -c 1 -o gpurun_out/prof_r1_v5a -f python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_v5a.log 2>&1
tail -2 gpurun_out/ncu_v5a.log

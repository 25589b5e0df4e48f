#!/bin/bash
mkdir -p gpurun_out
run() { r=$(env "$@" timeout 120 python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -E "value|rror" | python -c "import sys,json; [print(round(json.loads(l)['value'],1), round(json.loads(l)['ms_per_step'],4)) for l in sys.stdin]"); echo "$* :: $r"; }
{
for lz in 29 64; do
for dbg in 0 1 4 8 12 13; do
run GDM_FUSED_LZ=$lz GDM_FUSED_DBG=$dbg
done
done
} > gpurun_out/v4_ablate.log 2>&1
cat gpurun_out/v4_ablate.log
GDM_FUSED_LZ=29 timeout 300 ncu --set full --import-source on --clock-control none -k regex:kron3d -s 10 -c 1 -o gpurun_out/prof_r1_v4a -f python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_v4a.log 2>&1
tail -3 gpurun_out/ncu_v4a.log

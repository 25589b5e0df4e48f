#!/bin/bash
mkdir -p gpurun_out
for v in 0 4 7; do
  if [ $v = 0 ]; then export GDM_FUSED_FAMILY=3; unset GDM_FUSED_FAMILY; fi
  if [ $v = 4 ]; then unset GDM_FUSED_V4; export GDM_FUSED_FAMILY=4; fi
  if [ $v = 7 ]; then unset GDM_FUSED_V4; export GDM_FUSED_FAMILY=7; fi
  timeout 150 python tools/bench_ops.py --steps 30 2>&1 | grep fused | sed "s/^/family=$v /"
done > gpurun_out/ops_families.log 2>&1
cat gpurun_out/ops_families.log

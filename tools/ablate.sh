#!/bin/bash
for cfg in 14; do
for dbg in 0 1 4 8 16 32 48 17 49 12 28 29; do
  r=$(GDM_FUSED_LZ=29 GDM_FUSED_CFG=$cfg GDM_FUSED_DBG=$dbg python bench.py --steps 20 --warmup 5 --quick 2>&1 | tail -1)
  echo "cfg=$cfg dbg=$dbg $r"
done
done

#!/bin/bash
for cfg in 1 8; do
for dbg in 0 1 2 3 4 8 12 6 14 15; do
  r=$(GDM_FUSED_CFG=$cfg GDM_FUSED_DBG=$dbg python bench.py --steps 20 --warmup 5 --quick 2>&1 | tail -1)
  echo "cfg=$cfg dbg=$dbg $r"
done
done

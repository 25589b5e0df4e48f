#!/bin/bash
# sweep fused-kernel configurations on the BASELINE size (one process each; z-chunking self-tuned)
for cfg in ${CFGS:-14 6 12 17 18 19 20 11 1 3 10}; do
    r=$(GDM_FUSED_VERBOSE=1 GDM_FUSED_CFG=$cfg timeout 120 python bench.py --steps 20 --warmup 5 --quick 2>&1 | grep -E "gdm\]|value|rror" | tr '\n' ' ')
    echo "cfg=$cfg $r"
done

#!/bin/bash
# sweep fused-kernel configurations on the BASELINE size (one process each; z-chunking self-tuned)
for cfg in ${CFGS:-1 3 4 5 6 7 8 10 11 12 13 14}; do
    r=$(GDM_FUSED_VERBOSE=1 GDM_FUSED_CFG=$cfg python bench.py --steps 20 --warmup 5 --quick 2>&1 | grep -E "gdm\]|value" | tr '\n' ' ')
    echo "cfg=$cfg $r"
done

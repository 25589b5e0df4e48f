#!/bin/bash
mkdir -p gpurun_out
for cfg in 101 109 120 121 122 123 124 125 126; do
  r=$(GDM_FUSED_CFG=$cfg timeout 60 python tools/bench_ops.py --steps 30 --p 1 2>&1 | grep fused | python -c "import sys,json; print(' '.join(f\"{json.loads(l)['op'][:4]}={json.loads(l)['gdofs']}\" for l in sys.stdin))")
  echo "cfg=$cfg $r"
done > gpurun_out/p1_tune.log 2>&1
cat gpurun_out/p1_tune.log

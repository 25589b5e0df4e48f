#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x --timeout 180 > gpurun_out/pytest_gpu_q.log 2>&1; tail -3 gpurun_out/pytest_gpu_q.log
{
echo "fused passes:"; timeout 200 python tools/bench_ops.py --steps 20 2>&1 | grep generic
} > gpurun_out/generic_ops2.log 2>&1
cat gpurun_out/generic_ops2.log
bash tools/gpu_generic_prof.sh

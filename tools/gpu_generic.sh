#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x --timeout 180 > gpurun_out/pytest_gpu_q.log 2>&1; tail -3 gpurun_out/pytest_gpu_q.log
{
echo "fused passes (one launch per direction, marching window, flattened indices):"; timeout 200 python tools/bench_ops.py --steps 20 2>&1 | grep generic
echo "one-output passes (GDM_GENERIC_FUSED=0):"; GDM_GENERIC_FUSED=0 timeout 200 python tools/bench_ops.py --steps 20 2>&1 | grep generic
} > gpurun_out/generic_ops.log 2>&1
cat gpurun_out/generic_ops.log

#!/bin/bash
# round-2 session j: watchdog build (every wait bounded and recorded) on the guided self-scheduled partition that stalled in
# session h; sanity run of the same build on the static partition; fused-dot reproducibility with per-share partial sums
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WD=$PWD/dealii-galerkin-difference-methods_b200/libgdm_b200_wd.so
GDM_B200_LIB=$WD GDM_PERS_MODE=guided GDM_FUSED_VERBOSE=1 timeout 150 python bench.py --quick --steps 2 --warmup 1 > gpurun_out/j_wd_guided.log 2>&1
echo "rc=$?" >> gpurun_out/j_wd_guided.log
GDM_B200_LIB=$WD GDM_FUSED_VERBOSE=1 timeout 150 python bench.py --quick --steps 5 --warmup 3 > gpurun_out/j_wd_static.log 2>&1
echo "rc=$?" >> gpurun_out/j_wd_static.log
timeout 600 python -m pytest tests/test_gpu_fused.py -q > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j_pytest.log

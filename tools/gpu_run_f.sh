#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
GDM_PERS_CFG=831 timeout 600 python -m pytest tests/test_gpu_pers.py -x -q -k "seams and (stiffness or advection_t)" > gpurun_out/f_pytest_831.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest_831.log
run() {
  echo "=== $*" >> gpurun_out/f_bench.log
  env "$@" GDM_FUSED_VERBOSE=1 timeout 300 python bench.py --quick --steps 200 --warmup 20 >> gpurun_out/f_bench.log 2>&1
}
run GDM_PERS_CFG=825
run GDM_PERS_CFG=831
run GDM_PERS_CFG=832
run GDM_PERS_CFG=833
run GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1150,1150,1250
run GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1250,1250,1400
run GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1000,1000,1000
run GDM_PERS_CFG=833 GDM_PERS_WEIGHTS=1150,1150,1250
GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1000,1000,1000 GDM_PERS_TRACE=gpurun_out/f_trace831u.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 >> gpurun_out/f_trace.log 2>&1
GDM_PERS_CFG=833 GDM_PERS_WEIGHTS=1000,1000,1000 GDM_PERS_TRACE=gpurun_out/f_trace833u.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 >> gpurun_out/f_trace.log 2>&1

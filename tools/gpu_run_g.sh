#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
GDM_PERS_CFG=831 timeout 600 python -m pytest tests/test_gpu_pers.py -x -q -k "seams and (stiffness or advection_t)" > gpurun_out/g_pytest_831.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_pytest_831.log
run() {
  echo "=== $*" >> gpurun_out/g_bench.log
  env "$@" GDM_FUSED_VERBOSE=1 timeout 300 python bench.py --quick --steps 200 --warmup 20 >> gpurun_out/g_bench.log 2>&1
}
run GDM_PERS_CFG=825
run GDM_PERS_CFG=831
run GDM_PERS_CFG=833
run GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1150,1150,1250
run GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1250,1250,1400
run GDM_PERS_CFG=833 GDM_PERS_WEIGHTS=1150,1150,1250
run GDM_PERS_CFG=833 GDM_PERS_WEIGHTS=1250,1250,1400
run GDM_PERS_CFG=831 GDM_PERS_COSTS=4.0,0.0,0.0,0.0
run GDM_PERS_CFG=831 GDM_PERS_COSTS=4.0,1.2,2.0,0.3 GDM_PERS_WEIGHTS=1150,1150,1250
GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1000,1000,1000 GDM_PERS_TRACE=gpurun_out/g_trace831u.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 >> gpurun_out/g_trace.log 2>&1
GDM_PERS_CFG=833 GDM_PERS_WEIGHTS=1000,1000,1000 GDM_PERS_TRACE=gpurun_out/g_trace833u.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 >> gpurun_out/g_trace.log 2>&1
for p in 1 5; do timeout 300 python tools/bench_ops.py --steps 30 --p $p >> gpurun_out/g_ops.log 2>&1; done
GDM_PERS_CFG=835 timeout 300 python tools/bench_ops.py --steps 30 --p 1 >> gpurun_out/g_ops835.log 2>&1
GDM_PERS_CFG=836 timeout 300 python tools/bench_ops.py --steps 30 --p 5 >> gpurun_out/g_ops836.log 2>&1

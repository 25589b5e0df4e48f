#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in 831 830; do
GDM_PERS_CFG=$cfg timeout 900 python -m pytest tests/test_gpu_pers.py tests/test_gpu_fused.py -x -q -k "not baseline_size" > gpurun_out/e_pytest_$cfg.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest_$cfg.log
done
timeout 900 python -m pytest tests/test_gpu_pers.py tests/test_gpu_fused.py tests/test_gpu_apply.py tests/test_gpu_solvers.py -x -q -k "not baseline_size and not golden_counts" > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest.log
run() {
  echo "=== $*" >> gpurun_out/e_bench.log
  env "$@" GDM_FUSED_VERBOSE=1 timeout 300 python bench.py --quick --steps 200 --warmup 20 >> gpurun_out/e_bench.log 2>&1
}
run A=0
run GDM_PERS_CFG=821
run GDM_PERS_CFG=825
run GDM_PERS_CFG=830
run GDM_PERS_CFG=831
run GDM_PERS_CFG=832
run GDM_PERS_CFG=833
run GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1300,1300,1450
run GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1200,1200,1300
run GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1500,1500,1700
GDM_PERS_CFG=831 GDM_PERS_TRACE=gpurun_out/e_trace831.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 > gpurun_out/e_trace.log 2>&1
GDM_PERS_CFG=831 GDM_PERS_WEIGHTS=1000,1000,1000 GDM_PERS_TRACE=gpurun_out/e_trace831u.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 >> gpurun_out/e_trace.log 2>&1

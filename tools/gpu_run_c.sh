#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
run() {
  echo "=== $*" >> gpurun_out/c_bench.log
  env "$@" GDM_FUSED_VERBOSE=1 timeout 300 python bench.py --quick --steps 200 --warmup 20 >> gpurun_out/c_bench.log 2>&1
}
run A=0
run GDM_PERS_ALIGNED=0
run GDM_PERS_CFG=810
run GDM_PERS_CFG=811
run GDM_PERS_CFG=812
run GDM_PERS_CFG=814
run GDM_PERS_CFG=820
run GDM_PERS_CFG=821
run GDM_PERS_CFG=821 GDM_PERS_ALIGNED=0
run GDM_PERS_CFG=823
run GDM_PERS_CFG=824
run GDM_PERS_CFG=825
GDM_PERS_TRACE=gpurun_out/c_trace800.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 > gpurun_out/c_trace.log 2>&1
GDM_PERS_CFG=821 GDM_PERS_TRACE=gpurun_out/c_trace821.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 >> gpurun_out/c_trace.log 2>&1

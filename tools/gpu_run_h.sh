#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*" >> gpurun_out/h_bench.log
  env "$@" GDM_FUSED_VERBOSE=1 timeout 300 python bench.py --quick --steps 200 --warmup 20 >> gpurun_out/h_bench.log 2>&1
}
run A=0
run GDM_PERS_GUIDE=1.5,8
run GDM_PERS_GUIDE=0.75,8
run GDM_PERS_GUIDE=1.0,12
run GDM_PERS_GUIDE=1.0,6
run GDM_PERS_GUIDE=2.0,10
run GDM_PERS_MODE=static
GDM_PERS_TRACE=gpurun_out/h_trace.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 >> gpurun_out/h_trace.log 2>&1
timeout 600 python tools/bench_ops.py --steps 30 > gpurun_out/h_ops.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_pers.py tests/test_gpu_fused.py tests/test_gpu_apply.py tests/test_gpu_solvers.py -x -q > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log

#!/bin/bash
# round-2 session z: one-sided x/y rows from the constant bank (product build) vs from shared memory (libgdm_b200_prev.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=$PWD/dealii-galerkin-difference-methods_b200
run() {
  echo "=== $*" >> gpurun_out/z_bench.log
  env "$@" timeout 60 python bench.py --quick --steps 500 --warmup 20 >> gpurun_out/z_bench.log 2>&1
}
run A=0
run GDM_B200_LIB=$D/libgdm_b200_prev.so
run A=0
run GDM_B200_LIB=$D/libgdm_b200_prev.so
run GDM_PERS_WEIGHTS=1150,1150,1250
run GDM_PERS_WEIGHTS=1100,1100,1150
run GDM_PERS_WEIGHTS=1350,1350,1500
timeout 300 python tools/bench_ops.py --steps 30 > gpurun_out/z_ops.log 2>&1
timeout 300 python -m pytest tests/test_gpu_pers.py -q -x -k "seams and (stiffness or advection_t)" > gpurun_out/z_pytest.log 2>&1

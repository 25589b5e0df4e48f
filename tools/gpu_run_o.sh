#!/bin/bash
# round-2 session o: A/B on one box: share prefetch (product build) vs fetch at the share switch (libgdm_b200_np.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NP=$PWD/dealii-galerkin-difference-methods_b200/libgdm_b200_np.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/o_smi.log 2>&1
run() {
  echo "=== $*" >> gpurun_out/o_bench.log
  env "$@" timeout 100 python bench.py --quick --steps 300 --warmup 20 >> gpurun_out/o_bench.log 2>&1
}
run A=0
run GDM_B200_LIB=$NP
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,8
run GDM_B200_LIB=$NP GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,8
run A=0
run GDM_B200_LIB=$NP

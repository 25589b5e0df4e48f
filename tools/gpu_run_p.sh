#!/bin/bash
# round-2 session p: plane stream continues into the prefetched next share (no pipeline refill at a share switch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*" >> gpurun_out/p_bench.log
  env "$@" timeout 60 python bench.py --quick --steps 300 --warmup 20 >> gpurun_out/p_bench.log 2>&1
  echo "rc=$?" >> gpurun_out/p_bench.log
}
run A=0
run GDM_PERS_MODE=guided
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=1.5,8
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,8
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=3.0,8
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,12
timeout 300 python -m pytest tests/test_gpu_pers.py -x -q -k "seams and (stiffness or mass)" > gpurun_out/p_pytest_static.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest_static.log
GDM_PERS_MODE=guided timeout 300 python -m pytest tests/test_gpu_pers.py -x -q -k "seams and (stiffness or mass)" > gpurun_out/p_pytest_guided.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest_guided.log

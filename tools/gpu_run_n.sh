#!/bin/bash
# round-2 session n: next share (ticket, job range, first job) prefetched during the current one; static vs guided schedules
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*" >> gpurun_out/n_bench.log
  env "$@" timeout 100 python bench.py --quick --steps 300 --warmup 20 >> gpurun_out/n_bench.log 2>&1
}
run A=0
run GDM_PERS_MODE=guided
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=1.5,8
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,8
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=3.0,8
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=1.5,6
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,6
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,12
GDM_PERS_MODE=guided timeout 300 python -m pytest tests/test_gpu_pers.py -x -q -k "seams and (stiffness or mass)" > gpurun_out/n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/n_pytest.log

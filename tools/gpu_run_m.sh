#!/bin/bash
# round-2 session m: fast-body scratch planes; static vs guided level schedules; seam parity in guided mode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*" >> gpurun_out/m_bench.log
  env "$@" timeout 100 python bench.py --quick --steps 300 --warmup 20 >> gpurun_out/m_bench.log 2>&1
}
run A=0
run GDM_PERS_MODE=guided
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=0.75,8
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=1.5,8
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=1.0,12
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,8
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=0.5,8
GDM_PERS_MODE=guided timeout 400 python -m pytest tests/test_gpu_pers.py -x -q -k "seams" > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m_pytest.log

#!/bin/bash
# round-2 GPU session A: parity of the persistent kernel + variant timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
run() { # env... -- label
  echo "=== $*" >> gpurun_out/a_bench.log
  env "$@" GDM_FUSED_VERBOSE=1 timeout 300 python bench.py --quick --steps 200 --warmup 20 >> gpurun_out/a_bench.log 2>&1
}
run A=0
run GDM_PERS_ALIGNED=0
run GDM_PERS_SLOTS=592
run GDM_PERS_SLOTS=256
run GDM_PERS_SLOTS=148
run GDM_PERS_SLOTS=444
run GDM_PERS_CFG=810
run GDM_PERS_CFG=811
run GDM_PERS_CFG=812
run GDM_PERS_CFG=813
run GDM_PERS_CFG=814
run GDM_PERS_CFG=815
run GDM_FUSED_FAMILY=3
run GDM_FUSED_FAMILY=4
timeout 600 python tools/bench_ops.py --steps 30 > gpurun_out/a_ops.log 2>&1

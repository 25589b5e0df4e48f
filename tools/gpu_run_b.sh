#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*" >> gpurun_out/b_bench.log
  env "$@" GDM_FUSED_VERBOSE=1 timeout 300 python bench.py --quick --steps 200 --warmup 20 >> gpurun_out/b_bench.log 2>&1
}
run A=0
run GDM_PERS_ALIGNED=0
run GDM_PERS_CFG=812
run GDM_PERS_CFG=812 GDM_PERS_ALIGNED=0
run GDM_PERS_CFG=814
CMD="python bench.py --quick --steps 5 --warmup 3"
$CMD > gpurun_out/b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:kron3d_pers -s 4 -c 2 -o gpurun_out/prof_r2_p800 -f $CMD > gpurun_out/b_ncu.log 2>&1
GDM_PERS_CFG=812 $CMD > gpurun_out/b_plain812.log 2>&1 &&
GDM_PERS_CFG=812 ncu --set full --clock-control none --import-source on -k regex:kron3d_pers -s 4 -c 2 -o gpurun_out/prof_r2_p812 -f $CMD > gpurun_out/b_ncu812.log 2>&1

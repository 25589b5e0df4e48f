#!/bin/bash
# round-2 session z2: plane-cost weights of edge tiles for the build with constant-bank one-sided rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*" >> gpurun_out/z2_bench.log
  env "$@" timeout 60 python bench.py --quick --steps 500 --warmup 20 >> gpurun_out/z2_bench.log 2>&1
}
for w in 1300,1300,1450 1350,1350,1500 1400,1400,1550 1450,1450,1650 1550,1550,1750 1350,1250,1450 1250,1350,1450 1400,1300,1600 1300,1400,1600; do
run GDM_PERS_WEIGHTS=$w GDM_FUSED_VERBOSE=1
done

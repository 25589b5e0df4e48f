#!/bin/bash
# round-2 session s: how much x-pass work the issuer warp can carry (x-task cost model, c_issue) for both builds
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=$PWD/dealii-galerkin-difference-methods_b200
run() {
  echo "=== $*" >> gpurun_out/s_bench.log
  env "$@" timeout 60 python bench.py --quick --steps 300 --warmup 20 >> gpurun_out/s_bench.log 2>&1
  echo "rc=$?" >> gpurun_out/s_bench.log
}
for lib in libgdm_b200.so libgdm_b200_np.so; do
for ci in 0.3 1.0 2.0 4.0; do
run GDM_B200_LIB=$D/$lib GDM_PERS_COSTS=4.0,0.85,1.5,$ci
done
done
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,8 GDM_PERS_COSTS=4.0,0.85,1.5,2.0
run GDM_PERS_MODE=guided GDM_PERS_GUIDE=1.5,8 GDM_PERS_COSTS=4.0,0.85,1.5,2.0

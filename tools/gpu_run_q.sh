#!/bin/bash
# round-2 session q: bisect of the slowdown of session p (A/B builds on one box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=$PWD/dealii-galerkin-difference-methods_b200
run() {
  echo "=== $*" >> gpurun_out/q_bench.log
  env "$@" timeout 60 python bench.py --quick --steps 300 --warmup 20 >> gpurun_out/q_bench.log 2>&1
  echo "rc=$?" >> gpurun_out/q_bench.log
}
for lib in libgdm_b200.so libgdm_b200_np.so libgdm_b200_ns.so libgdm_b200_nt.so libgdm_b200_nn.so; do
run GDM_B200_LIB=$D/$lib
run GDM_B200_LIB=$D/$lib GDM_PERS_MODE=guided GDM_PERS_GUIDE=2.0,8
done

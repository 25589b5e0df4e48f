#!/bin/bash
# round-2 session aa: measured plan search at operator creation vs the fixed default weights
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*" >> gpurun_out/aa_bench.log
  env "$@" timeout 60 python bench.py --quick --steps 500 --warmup 20 >> gpurun_out/aa_bench.log 2>&1
}
run GDM_FUSED_VERBOSE=1
run GDM_PERS_TUNE=0
run GDM_FUSED_VERBOSE=1
run GDM_PERS_TUNE=0
timeout 300 python tools/bench_ops.py --steps 30 > gpurun_out/aa_ops.log 2>&1

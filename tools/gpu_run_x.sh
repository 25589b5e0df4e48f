#!/bin/bash
# round-2 session x: face kernel behind the tile kernel on a lowest-priority stream (A/B by switch), PCIe rates of the box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*" >> gpurun_out/x_bench.log
  env "$@" timeout 60 python bench.py --quick --steps 500 --warmup 20 >> gpurun_out/x_bench.log 2>&1
}
run A=0
run GDM_FACE_ORDER=first
run A=0
run GDM_FACE_ORDER=first
timeout 120 python tools/pcie_probe.py > gpurun_out/x_pcie.json 2>&1
timeout 400 python bench.py --steps 500 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_apply.py -q -x > gpurun_out/x_pytest.log 2>&1

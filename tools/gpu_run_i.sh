#!/bin/bash
# round-2 session i: full GPU suite on the static default, driver-style bench line, launch list, ncu capture of cfg 833
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/i_smi.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest.log
timeout 600 python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
CMD="python bench.py --quick --steps 20 --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/i_launches.csv $CMD > gpurun_out/i_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kron3d_pers -s 6 -c 1 -o /tmp/prof_i833 -f $CMD > gpurun_out/i_ncu.log 2>&1
ncu -i /tmp/prof_i833.ncu-rep --page raw --csv > gpurun_out/i833_raw.csv 2>/dev/null
ncu -i /tmp/prof_i833.ncu-rep --page source --csv > gpurun_out/i833_source.csv 2>/dev/null
ls -la /tmp/prof_i833.ncu-rep >> gpurun_out/i_ncu.log
timeout 600 python tools/bench_ops.py --steps 30 > gpurun_out/i_ops.log 2>&1

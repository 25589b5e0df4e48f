#!/bin/bash
# round-2 session ab: final state -- full GPU suite, driver-style bench lines, launch list, ncu capture, operator table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/ab_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ab_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ab_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err
CMD="python bench.py --quick --steps 20 --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ab_launches.csv $CMD > gpurun_out/ab_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kron3d_pers -s 70 -c 1 -o /tmp/prof_ab -f $CMD > gpurun_out/ab_ncu.log 2>&1
ncu -i /tmp/prof_ab.ncu-rep --page raw --csv > gpurun_out/ab_raw.csv 2>/dev/null
ncu -i /tmp/prof_ab.ncu-rep --page source --csv > gpurun_out/ab_source.csv 2>/dev/null
timeout 600 python tools/bench_ops.py --steps 30 > gpurun_out/ab_ops.log 2>&1

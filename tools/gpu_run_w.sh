#!/bin/bash
# round-2 session w: final state -- full GPU suite, driver-style bench lines (ours + reference arm), launch list, ncu capture,
# operator table, RK workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/w_smi.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/w_bench_ref.json 2> gpurun_out/w_bench_ref.err
CMD="python bench.py --quick --steps 20 --warmup 3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/w_launches.csv $CMD > gpurun_out/w_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kron3d_pers -s 6 -c 1 -o /tmp/prof_w -f $CMD > gpurun_out/w_ncu.log 2>&1
ncu -i /tmp/prof_w.ncu-rep --page raw --csv > gpurun_out/w_raw.csv 2>/dev/null
ncu -i /tmp/prof_w.ncu-rep --page source --csv > gpurun_out/w_source.csv 2>/dev/null
timeout 600 python tools/bench_ops.py --steps 30 > gpurun_out/w_ops.log 2>&1
timeout 300 python bench.py --workload wave_rk4 --cells 192 --steps 5 > gpurun_out/w_wave.json 2> gpurun_out/w_wave.err
timeout 300 python bench.py --workload advection_rk4 --cells 192 --steps 5 > gpurun_out/w_adv.json 2> gpurun_out/w_adv.err

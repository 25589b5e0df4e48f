#!/bin/bash
# round-1 final measurements on one B200: GPU tests, bench line, reference arm, launch list under ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/pytest_gpu_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_final.json 2>&1; tail -c 300 gpurun_out/bench_ref_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 20 --warmup 3 --cg-steps 10 > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/launches_final.csv
timeout 200 python tools/bench_ops.py --steps 30 > gpurun_out/ops_final.jsonl 2>&1; grep -c fused gpurun_out/ops_final.jsonl

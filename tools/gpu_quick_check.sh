#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x --timeout 180 > gpurun_out/pytest_gpu_q.log 2>&1; tail -6 gpurun_out/pytest_gpu_q.log
timeout 300 python bench.py --steps 300 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, round(d['roofline']['frac'],4), d['e2e']['value'], d['cg'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
GDM_CG_FUSED_DOT=0 timeout 300 python bench.py --steps 100 > gpurun_out/bench_q2.json 2> gpurun_out/bench_q2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q2.json'))
print("separate dot:", d['cg'])
PY

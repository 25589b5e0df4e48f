#!/bin/bash
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_8gpu.json'))
print({k:d[k] for k in ['value','ms_per_step','n_gpus']}, round(d['roofline']['frac'],4), d['e2e']['value'], d['cg'])
PY

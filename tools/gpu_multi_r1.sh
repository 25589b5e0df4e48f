#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 250 > gpurun_out/pytest_multi_$N.log 2>&1; tail -3 gpurun_out/pytest_multi_$N.log
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print({k:d[k] for k in ['value','ms_per_step','n_gpus']}, round(d['roofline']['frac'],4), d['e2e']['value'], d['cg'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 | tail -c 300

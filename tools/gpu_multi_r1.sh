#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 200 > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 700 gpurun_out/bench_2gpu.json

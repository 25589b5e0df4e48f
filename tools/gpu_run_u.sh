#!/bin/bash
# round-2 session u (2 GPUs): multi-GPU parity (incl. pipelined host-buffer apply) and the 2-GPU bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/u_pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/u_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 10 > gpurun_out/u_bench_2gpu.json 2> gpurun_out/u_bench_2gpu.err
echo "rc=$?" >> gpurun_out/u_bench_2gpu.err

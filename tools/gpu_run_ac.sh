#!/bin/bash
# round-2 session ac (4 GPUs): parity with interior ranks (two neighbours) and the 4-GPU bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/ac_pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ac_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 200 --warmup 10 --cg-steps 20 > gpurun_out/ac_bench_4gpu.json 2> gpurun_out/ac_bench_4gpu.err
echo "rc=$?" >> gpurun_out/ac_bench_4gpu.err

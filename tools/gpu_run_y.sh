#!/bin/bash
# round-2 session y (2 GPUs): multi-GPU parity and bench line after the schedule changes (exchange first, face kernel on the
# lowest-priority stream during the exchange, finer host-pipeline chunks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/y_pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/y_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 10 > gpurun_out/y_bench_2gpu.json 2> gpurun_out/y_bench_2gpu.err
echo "rc=$?" >> gpurun_out/y_bench_2gpu.err
timeout 300 python bench.py --steps 200 --cg-steps 0 > gpurun_out/y_bench_1gpu.json 2> gpurun_out/y_bench_1gpu.err

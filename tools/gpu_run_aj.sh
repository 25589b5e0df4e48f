#!/bin/bash
# round-2 session aj: the driver's bench command on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python bench.py --steps 300 > gpurun_out/aj_bench.json 2> gpurun_out/aj_bench.err
echo "rc=$?" >> gpurun_out/aj_bench.err

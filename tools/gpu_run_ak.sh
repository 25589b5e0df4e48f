#!/bin/bash
# round-2 session ak: chunk count of the host-buffer pipeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in default 4 6 16 24; do
  if [ $c = default ]; then timeout 60 python tools/bench_e2e.py >> gpurun_out/ak_e2e.log 2>&1; else GDM_HOST_CHUNKS=$c timeout 60 python tools/bench_e2e.py >> gpurun_out/ak_e2e.log 2>&1; fi
done

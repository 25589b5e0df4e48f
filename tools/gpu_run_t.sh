#!/bin/bash
# round-2 session t: ncu capture of the build with the streamed share switch (25 % slower than the build without; why?)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --quick --steps 10 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kron3d_pers -s 6 -c 1 -o /tmp/prof_t -f $CMD > gpurun_out/t_ncu.log 2>&1
ncu -i /tmp/prof_t.ncu-rep --page raw --csv > gpurun_out/t_raw.csv 2>/dev/null
ncu -i /tmp/prof_t.ncu-rep --page source --csv > gpurun_out/t_source.csv 2>/dev/null

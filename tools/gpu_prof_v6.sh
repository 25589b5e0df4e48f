#!/bin/bash
mkdir -p gpurun_out
CFG=${CFG:-304}
GDM_FUSED_CFG=$CFG GDM_FUSED_LZ=${LZ:-43} timeout 300 ncu --set full --import-source on --clock-control none -k regex:kron3d -s 10 -c 1 -o /tmp/prof_v6 -f python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_v6a.log 2>&1
tail -2 gpurun_out/ncu_v6a.log
ncu -i /tmp/prof_v6.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_v6_${CFG}_source.csv 2>/dev/null
ncu -i /tmp/prof_v6.ncu-rep --page raw --csv > gpurun_out/prof_v6_${CFG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_v6.ncu-rep --page details --csv > gpurun_out/prof_v6_${CFG}_details.csv 2>/dev/null
ls -la gpurun_out /tmp/prof_v6.ncu-rep

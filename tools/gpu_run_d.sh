#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log
run() {
  echo "=== $*" >> gpurun_out/d_bench.log
  env "$@" GDM_FUSED_VERBOSE=1 timeout 300 python bench.py --quick --steps 200 --warmup 20 >> gpurun_out/d_bench.log 2>&1
}
run A=0
run GDM_PERS_ALIGNED=0
run GDM_PERS_WEIGHTS=1000,1000,1000
run GDM_PERS_WEIGHTS=1300,1250,1400
run GDM_PERS_WEIGHTS=1500,1450,1650
run GDM_PERS_CFG=821
run GDM_PERS_CFG=821 GDM_PERS_WEIGHTS=1410,1450,1580
run GDM_PERS_CFG=825 GDM_PERS_WEIGHTS=1410,1450,1580
run GDM_PERS_CFG=812
run GDM_PERS_CFG=814
GDM_PERS_TRACE=gpurun_out/d_trace800.txt timeout 120 python bench.py --quick --steps 3 --warmup 3 > gpurun_out/d_trace.log 2>&1
CMD="python bench.py --quick --steps 5 --warmup 3"
$CMD > gpurun_out/d_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:kron3d_pers -s 4 -c 1 -o /tmp/prof_r2_d800 -f $CMD > gpurun_out/d_ncu.log 2>&1
ncu -i /tmp/prof_r2_d800.ncu-rep --page raw --csv > gpurun_out/d800_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2_d800.ncu-rep --page source --csv > gpurun_out/d800_source.csv 2>/dev/null
ls -la /tmp/prof_r2_d800.ncu-rep >> gpurun_out/d_ncu.log
SZ=$(stat -c %s /tmp/prof_r2_d800.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 0 ] && [ "$SZ" -lt 40000000 ]; then cp /tmp/prof_r2_d800.ncu-rep gpurun_out/; fi
du -sh gpurun_out >> gpurun_out/d_ncu.log

#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'win_kernel' -o gpurun_out/dw5 -f \
   python scripts/profile_ops.py --set effi --only dw5_240_28,dw5s2_144 --reps 1 --warm 0 > gpurun_out/ncu_dw5.log 2>&1
grep -v "^==" gpurun_out/ncu_dw5.log | tail -4
ncu -i gpurun_out/dw5.ncu-rep --page raw --csv > gpurun_out/dw5_raw.csv 2>/dev/null
ncu -i gpurun_out/dw5.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/dw5_source.csv.gz
rm -f gpurun_out/*.ncu-rep
python scripts/ncu_summary.py gpurun_out/dw5_raw.csv | cut -c1-900

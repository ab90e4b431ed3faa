#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'win_kernel' -o gpurun_out/win2 -f python scripts/profile_ops.py --set mobilenet --reps 1 --warm 0 --only dw3_144 > gpurun_out/ncu_win2.log 2>&1
ncu -i gpurun_out/win2.ncu-rep --page raw --csv > gpurun_out/win2_raw.csv 2>/dev/null
ncu -i gpurun_out/win2.ncu-rep --page source --csv > gpurun_out/win2_source.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'igemm' -o gpurun_out/halo1 -f python scripts/profile_ops.py --set resnet50 --reps 1 --warm 0 --only c3_ > gpurun_out/ncu_halo1.log 2>&1
ncu -i gpurun_out/halo1.ncu-rep --page raw --csv > gpurun_out/halo1_raw.csv 2>/dev/null
ncu -i gpurun_out/halo1.ncu-rep --page source --csv > gpurun_out/halo1_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep; gzip -f gpurun_out/*_source.csv
ls -la gpurun_out

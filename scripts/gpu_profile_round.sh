#!/bin/bash
# Evidence for profiles/: (1) launch list of one bench run (shares of the step), (2) ncu --set full of the dominant
# kernel (the stem conv = first igemm2 launch of a forward) and of one kernel per family, exported to CSV on the box.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 140 --csv --log-file gpurun_out/launches_resnet50.csv \
   python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm2_kernel -c 1 -o gpurun_out/stem -f \
   python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_stem.log 2>&1
ncu -i gpurun_out/stem.ncu-rep --page raw --csv > gpurun_out/stem_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:'igemm|win_kernel|gap_kernel' -o gpurun_out/fam -f \
   python scripts/profile_ops.py --set resnet50,mobilenet,pool --reps 1 --warm 0 > gpurun_out/ncu_fam.log 2>&1
ncu -i gpurun_out/fam.ncu-rep --page raw --csv > gpurun_out/fam_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | head -30

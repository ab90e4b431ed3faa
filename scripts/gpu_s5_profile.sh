#!/bin/bash
# Evidence for profiles/ (session 5): launch list of one bench run, ncu --set full of the stem+pool kernel inside the
# bench and of the dominant / representative conv kernels in isolation; CSV exports made on the box.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_resnet50_s5.csv \
   python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_halo_kernel -s 2 -c 1 -o gpurun_out/stem_s5 -f \
   python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_stem.log 2>&1
ncu -i gpurun_out/stem_s5.ncu-rep --page raw --csv > gpurun_out/stem_s5_raw.csv 2>/dev/null
ncu -i gpurun_out/stem_s5.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/stem_s5_source.csv.gz
timeout 400 ncu --set full --clock-control none -k regex:'igemm2_kernel|igemm3_kernel' -o gpurun_out/ops_s5 -f \
   python scripts/profile_ops.py --set resnet50 --only c1_64_256_56_res,c3_64_56,c3_128_28,c3_256_14,c1_1024_256_14,c1_256_64_56 --reps 1 --warm 0 > gpurun_out/ncu_ops.log 2>&1
cat gpurun_out/ncu_ops.log | grep -v "^==" | tail -8
ncu -i gpurun_out/ops_s5.ncu-rep --page raw --csv > gpurun_out/ops_s5_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/*s5* 

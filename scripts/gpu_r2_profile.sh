#!/bin/bash
# Round-2 evidence for profiles/: launch list of one eager bench step (shares), ncu --set full of (a) the dominant kernel
# family of the ResNet-50 step in isolation, (b) the fp32 tier's tensor-core kernel inside the ResNet-18 bench, (c) the
# fused bottleneck-tail kernel inside the ResNet-50 bench.  CSV exports are made on the box; .ncu-rep files are dropped.
mkdir -p gpurun_out
B="--no-cpu-baseline --no-configs --graph 0 --sustain-s 0.01"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench_resnet50.csv \
   python bench.py --steps 2 --warmup 3 $B > gpurun_out/r02_ncu_launch.log 2>&1
tail -1 gpurun_out/r02_ncu_launch.log | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench_resnet18_fp32.csv \
   python bench.py --model resnet18 --steps 2 --warmup 3 $B > gpurun_out/r02_ncu_launch18.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:'igemm2_kernel|igemm3_kernel' -o gpurun_out/r02_ops -f \
   python scripts/profile_ops.py --set resnet50 --only c1_64_256_56_res,c3_64_56,c3_128_28,c3_256_14,c3_512_7,c1_2048_512_7,c1_256_1024_14_res --reps 1 --warm 0 > gpurun_out/r02_ncu_ops.log 2>&1
grep -v "^==" gpurun_out/r02_ncu_ops.log | tail -8
ncu -i gpurun_out/r02_ops.ncu-rep --page raw --csv > gpurun_out/r02_ops_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:f32x3_kernel -s 25 -c 8 -o gpurun_out/r02_f32x3 -f \
   python bench.py --model resnet18 --steps 1 --warmup 3 $B > gpurun_out/r02_ncu_f32x3.log 2>&1
ncu -i gpurun_out/r02_f32x3.ncu-rep --page raw --csv > gpurun_out/r02_f32x3_raw.csv 2>/dev/null
PCV_FUSE_TAIL=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm3x_kernel -s 3 -c 1 -o gpurun_out/r02_fused -f \
   python bench.py --steps 1 --warmup 3 $B > gpurun_out/r02_ncu_fused.log 2>&1
ncu -i gpurun_out/r02_fused.ncu-rep --page raw --csv > gpurun_out/r02_fused_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r02_*raw.csv gpurun_out/r02_launches*.csv

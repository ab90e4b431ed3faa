#!/bin/bash
# final build with the dual-source conv: the one test that changed, the ResNet-50 launch list (ncu gpu__time_duration.sum),
# and one ncu --set full capture of the dual-source instantiation (stage 2: 128->512 @28x28 + 1x1 s2 256@56x56)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nets.py -q -x -m gpu -k "fused_bottleneck_tail_whole_net or projection_shortcut" 2>&1 | tail -3
B="--no-cpu-baseline --no-configs --graph 0 --sustain-s 0.01"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_final_resnet50.csv \
   python bench.py --model resnet50 --steps 2 --warmup 3 $B > gpurun_out/r02_ncu_launch_resnet50.log 2>&1
tail -1 gpurun_out/r02_ncu_launch_resnet50.log | cut -c1-160
# launches 0..3 of the DUAL instantiation in a forward are stages 1..4; skip the compile-time / warm-up forwards' first one
timeout 400 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'igemm2_kernel<256, 4, false, true>|igemm2_kernel<256, 4, 0, 1>|igemm2_kernel<\(int\)256, \(int\)4, \(bool\)0, \(bool\)1>' -s 4 -c 4 -o gpurun_out/r02e_dual -f \
   python bench.py --model resnet50 --steps 1 --warmup 3 $B > gpurun_out/r02e_ncu_dual.log 2>&1
tail -1 gpurun_out/r02e_ncu_dual.log | cut -c1-120
ncu -i gpurun_out/r02e_dual.ncu-rep --page raw --csv > gpurun_out/r02e_dual_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r02e_dual_raw.csv gpurun_out/r02_launches_final_resnet50.csv

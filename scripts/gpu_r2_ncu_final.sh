#!/bin/bash
# ncu --set full of the dominant kernels of configs 3 and 4 in the final build: the fused expansion -> dw -> pw kernel
# (16 -> 96 -> 24 @112x112, stride 2) and the gated conv3 (128 -> 256 @56x56): DRAM traffic against the fused groups' algorithmic bytes
mkdir -p gpurun_out
B="--no-cpu-baseline --no-configs --graph 0 --sustain-s 0.01"
timeout 400 ncu --set full --clock-control none -k regex:xdwpw_kernel -s 3 -c 1 -o gpurun_out/r02d_xdwpw -f \
   python bench.py --model mobilenetv2_w1 --steps 1 --warmup 3 $B > gpurun_out/r02d_ncu_xdwpw.log 2>&1
tail -1 gpurun_out/r02d_ncu_xdwpw.log | cut -c1-120
ncu -i gpurun_out/r02d_xdwpw.ncu-rep --page raw --csv > gpurun_out/r02d_xdwpw_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'igemm2_kernel<256, 4, true>|igemm2_kernel<256, 4, 1>|igemm2_kernel<\(int\)256, \(int\)4, \(bool\)1>' -s 16 -c 2 -o gpurun_out/r02d_gate -f \
   python bench.py --model seresnext50_32x4d --steps 1 --warmup 3 $B > gpurun_out/r02d_ncu_gate.log 2>&1
tail -1 gpurun_out/r02d_ncu_gate.log | cut -c1-120
ncu -i gpurun_out/r02d_gate.ncu-rep --page raw --csv > gpurun_out/r02d_gate_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r02d_*raw.csv

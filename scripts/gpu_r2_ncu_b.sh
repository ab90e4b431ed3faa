#!/bin/bash
# Round-2 (session 3) evidence: ncu --set full of one eager MobileNetV2 fp16 step (fused dw->pw, expansions, depthwise) and
# of the SE-ResNeXt-50 step's grouped / squeeze / excite kernels.  CSV exports are made on the box; .ncu-rep dropped.
mkdir -p gpurun_out
B="--no-cpu-baseline --no-configs --graph 0 --sustain-s 0.01"
timeout 600 ncu --set full --clock-control none -k regex:'dwpw|win_kernel|igemm2_kernel|igemm_kernel|stem_halo' -s 90 -c 45 -o gpurun_out/r02b_mnv2 -f \
   python bench.py --model mobilenetv2_w1 --steps 1 --warmup 3 $B > gpurun_out/r02b_ncu_mnv2.log 2>&1
tail -2 gpurun_out/r02b_ncu_mnv2.log | cut -c1-300
ncu -i gpurun_out/r02b_mnv2.ncu-rep --page raw --csv > gpurun_out/r02b_mnv2_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:'igemm3_kernel|gap_kernel|se_excite|se_fc|igemm2_kernel<64' -s 130 -c 65 -o gpurun_out/r02b_sex -f \
   python bench.py --model seresnext50_32x4d --steps 1 --warmup 3 $B > gpurun_out/r02b_ncu_sex.log 2>&1
tail -2 gpurun_out/r02b_ncu_sex.log | cut -c1-300
ncu -i gpurun_out/r02b_sex.ncu-rep --page raw --csv > gpurun_out/r02b_sex_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r02b_*raw.csv

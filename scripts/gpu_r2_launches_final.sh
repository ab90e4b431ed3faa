#!/bin/bash
# launch lists (ncu gpu__time_duration.sum, --clock-control none) of one eager bench step of the final build: ResNet-50,
# MobileNetV2 (fused dw->pw / expansion->dw->pw kernels) and SE-ResNeXt-50 (gated conv3 epilogue)
mkdir -p gpurun_out
B="--no-cpu-baseline --no-configs --graph 0 --sustain-s 0.01"
for m in resnet50 mobilenetv2_w1 seresnext50_32x4d; do
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_final_$m.csv \
   python bench.py --model $m --steps 2 --warmup 3 $B > gpurun_out/r02_ncu_launch_$m.log 2>&1
tail -1 gpurun_out/r02_ncu_launch_$m.log | cut -c1-160
done
ls -la gpurun_out/r02_launches_final_*.csv

#!/bin/bash
# source-level ncu capture of the fused dw->pw kernel (C=144 @56x56 +res, the third dwpw launch of a MobileNetV2 forward):
# where do the stencil warps stall?  Exports the SASS page with per-instruction stall samples.
mkdir -p gpurun_out
B="--no-cpu-baseline --no-configs --graph 0 --sustain-s 0.01"
PCV_FUSE_XDWPW=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'^dwpw_kernel' -s 2 -c 1 -o gpurun_out/r02c_dwpw -f \
   python bench.py --model mobilenetv2_w1 --steps 1 --warmup 3 $B > gpurun_out/r02c_ncu_dwpw.log 2>&1
tail -2 gpurun_out/r02c_ncu_dwpw.log | cut -c1-200
ncu -i gpurun_out/r02c_dwpw.ncu-rep --page source --csv > gpurun_out/r02c_dwpw_source.csv 2>/dev/null
ncu -i gpurun_out/r02c_dwpw.ncu-rep --page raw --csv > gpurun_out/r02c_dwpw_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r02c_*

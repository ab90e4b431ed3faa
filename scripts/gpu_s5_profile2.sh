#!/bin/bash
# Evidence for profiles/ (session 5, part 2): launch lists of the other configs + ncu --set full of the bandwidth kernels
mkdir -p gpurun_out
for m in mobilenetv2_w1 seresnext50_32x4d deeplabv3_resnetd50b_voc efficientnet_b0; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${m}_s5.csv \
     python bench.py --model $m --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_launch_$m.log 2>&1
  tail -1 gpurun_out/ncu_launch_$m.log | cut -c1-160
done
timeout 400 ncu --set full --clock-control none -k regex:'win_kernel|gap_kernel|se_scale_kernel|s2d_ingest2|bilinear_nchw_rows' -c 14 -o gpurun_out/bw_s5 -f \
   python scripts/profile_ops.py --set mobilenet,pool --only dw3_32_112,dw3s2_96_112,dw3_144_56,dw3_384_14,maxpool_64_112,gavg_256_3136 --reps 1 --warm 0 > gpurun_out/ncu_bw.log 2>&1
grep -v "^==" gpurun_out/ncu_bw.log | tail -8
ncu -i gpurun_out/bw_s5.ncu-rep --page raw --csv > gpurun_out/bw_s5_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
python scripts/ncu_summary.py gpurun_out/bw_s5_raw.csv | cut -c1-420
ls -la gpurun_out/launches_*_s5.csv

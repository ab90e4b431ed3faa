#!/bin/bash
# ncu --set full of the dominant kernels of the other configs (for profiles/ncu_traffic.json)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none -k regex:'igemm2_kernel|se_scale_kernel|gap_kernel' -c 8 -o gpurun_out/dom_s5 -f \
   python scripts/profile_ops.py --set deeplab,mobilenet,pool --only c3_d12_2048,c1_16_96_112,sescale_256,gavg_256_3136 --reps 1 --warm 0 > gpurun_out/ncu_dom.log 2>&1
grep -v "^==" gpurun_out/ncu_dom.log | tail -6
ncu -i gpurun_out/dom_s5.ncu-rep --page raw --csv > gpurun_out/dom_s5_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
python scripts/ncu_summary.py gpurun_out/dom_s5_raw.csv | cut -c1-330

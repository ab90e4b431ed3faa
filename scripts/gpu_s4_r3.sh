#!/bin/bash
# ncu --set full of the stem kernel (source-level) inside one bench forward
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stem_halo -c 1 -o gpurun_out/stemh -f \
   python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_stemh.log 2>&1
tail -3 gpurun_out/ncu_stemh.log
ncu -i gpurun_out/stemh.ncu-rep --page raw --csv > gpurun_out/stemh_raw.csv 2>/dev/null
ncu -i gpurun_out/stemh.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/stemh_source.csv.gz
python scripts/ncu_summary.py gpurun_out/stemh_raw.csv
python scripts/ncu_source_mix.py gpurun_out/stemh_source.csv.gz
rm -f gpurun_out/*.ncu-rep

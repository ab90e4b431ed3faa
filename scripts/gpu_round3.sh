#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dw or maxpool or se_excite or bilinear" > gpurun_out/pytest_kernels.log 2>&1; tail -5 gpurun_out/pytest_kernels.log
timeout 120 python scripts/profile_ops.py --set mobilenet,pool > gpurun_out/profile_ops2.log 2>&1; cat gpurun_out/profile_ops2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'igemm' -o gpurun_out/igemm_full -f python scripts/profile_ops.py --set resnet50 --reps 1 --warm 0 > gpurun_out/ncu_igemm.log 2>&1
ncu -i gpurun_out/igemm_full.ncu-rep --page raw --csv > gpurun_out/igemm_full_raw.csv 2>/dev/null
ncu -i gpurun_out/igemm_full.ncu-rep --page source --csv > gpurun_out/igemm_full_source.csv 2>/dev/null
ls -la gpurun_out/
sz=$(stat -c %s gpurun_out/igemm_full.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm gpurun_out/igemm_full.ncu-rep; fi
sz=$(stat -c %s gpurun_out/igemm_full_source.csv); if [ "$sz" -gt 25000000 ]; then gzip gpurun_out/igemm_full_source.csv; fi
ls -la gpurun_out/

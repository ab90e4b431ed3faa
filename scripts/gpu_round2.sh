#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_kernels.log 2>&1; tail -15 gpurun_out/pytest_kernels.log
timeout 120 python scripts/profile_ops.py --set mobilenet,pool > gpurun_out/profile_ops2.log 2>&1; cat gpurun_out/profile_ops2.log
for m in mobilenetv2_w1 seresnext50_32x4d deeplabv3_resnetd50b_voc; do
  timeout 150 python bench.py --model $m --no-cpu-baseline --ops-out gpurun_out/bench_ops_$m.json > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$m.json").read().strip().splitlines()[-1]); print("$m", d["value"], d["ms_per_step"], d["roofline_step"])
except Exception as e: print("$m failed", e); print(open("gpurun_out/bench_$m.err").read()[-1500:])
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'win_kernel' -o gpurun_out/win_full -f python scripts/profile_ops.py --set mobilenet,pool --reps 1 --warm 0 --only _1 > gpurun_out/ncu_win.log 2>&1
ncu -i gpurun_out/win_full.ncu-rep --page raw --csv > gpurun_out/win_full_raw.csv 2>/dev/null
ncu -i gpurun_out/win_full.ncu-rep --page source --csv > gpurun_out/win_full_source.csv 2>/dev/null
ls -la gpurun_out/
